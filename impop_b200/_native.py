"""ctypes binding of libimpop_b200.so -- the C ABI declared in include/impop_b200.h.

There is no fallback of any kind: if the shared library cannot be loaded (and cannot be
built because nvcc is absent) importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

NSTATS = 20
NCOUNTS = 8
LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = 1, 2, 4, 8
ALGO_TCGEN05, ALGO_SIMT = 0, 1
COMPACT_PAIRS, COMPACT_REPLICATE = 1, 2
KERNELS = {"prep": 0, "pairs": 1, "sums": 2, "colstat": 3, "finalize": 4, "sites": 5}
ST = {"pi": 0, "pi_per_site": 1, "pi_a": 2, "pi_b": 3, "pi_xy": 4, "dxy": 5, "da": 6, "fst": 7, "S": 8,
      "tajima_d": 9, "a1": 10, "e1": 11, "e2": 12, "n": 13, "sum_S": 14, "sum_AA": 15, "sum_BB": 16,
      "sum_AB": 17, "tajima_d_raw": 18, "S_bubbles": 19}

ERRORS = {-1: "IMPOP_ERR_ARG", -2: "IMPOP_ERR_CUDA", -3: "IMPOP_ERR_NOMEM", -4: "IMPOP_ERR_RANGE", -5: "IMPOP_ERR_DEVICE"}

_p = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double


class BatchDesc(C.Structure):
    """impop_batch_desc_t"""
    _fields_ = [("windows", _i32), ("n_host", _p), ("m_host", _p), ("pitch_words_host", _p), ("x_off_host", _p),
                ("len_off_host", _p), ("lab_off_host", _p), ("length_host", _p), ("x_dev", _p), ("node_len_dev", _p),
                ("labels_dev", _p), ("node_len_host", _p), ("stream", _p), ("site_runs_host", _p),
                ("row_adj_dev", _p), ("win_const_host", _p), ("col_mult_dev", _p), ("heavy_entries_host", _p)]


class GfaInfo(C.Structure):
    """impop_gfa_info_t"""
    _fields_ = [("segments", _i64), ("paths", _i64), ("name_bytes", _i64), ("steps", _i64), ("error_line", _i64)]


class TsvInfo(C.Structure):
    """impop_tsv_info_t"""
    _fields_ = [("rows", _i64), ("names", _i64), ("name_bytes", _i64), ("status", _i32), ("reserved", _i32)]


# name -> (restype, argtypes); every symbol include/impop_b200.h declares
SIGNATURES = {
    "impop_version": (C.c_int, []),
    "impop_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "impop_destroy": (C.c_int, [_p]),
    "impop_last_error": (C.c_char_p, [_p]),
    "impop_check": (C.c_int, [_p, _p]),
    "impop_launch_count": (_i64, [_p]),
    "impop_timing_enable": (C.c_int, [_p, _i32]),
    "impop_timing_read": (C.c_int, [_p, _i32, C.POINTER(_f64), C.POINTER(_i64)]),
    "impop_pack_bits": (C.c_int, [_p, _p, _i32, _i32, _i64, _p, _i32, _p]),
    "impop_batch_create": (C.c_int, [_p, C.POINTER(BatchDesc), C.POINTER(_p)]),
    "impop_batch_destroy": (C.c_int, [_p, _p]),
    "impop_batch_items": (_i64, [_p]),
    "impop_window_stats": (C.c_int, [_p, _p, _i32, _p, _p, _p]),
    "impop_window_sums": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "impop_window_finalize": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "impop_pairwise": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p]),
    "impop_reduce_identity": (C.c_int, [_p, _p, _i32, _i64, _p, _p, _i64, _f64, _p, _p, _p, _p]),
    "impop_tajima_d": (C.c_int, [_p, _p, _p, _p, _i32, _p, _p, _p]),
    "impop_site_counts": (C.c_int, [_p, _p, _i64, _i32, _p, _i32, _p, _p, _p]),
    "impop_cluster": (C.c_int, [_p, _p, _i32, _i64, _f64, _p, _p]),
    "impop_selftest_division": (C.c_int, [_p, C.c_uint64, _i64, C.POINTER(_i64), _p]),
    "impop_debug_role_times": (C.c_int, [_p, _p, _i32]),
    "impop_greedy_groups": (C.c_int, [_p, _p, _i32, _i64, _f64, _p, _p, _p]),
    "impop_round_decimal": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "impop_repitch_rows": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "impop_dev_alloc": (C.c_int, [_p, _i64, C.POINTER(_p)]),
    "impop_dev_free": (C.c_int, [_p, _p]),
    "impop_dev_copy": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "impop_tsv_scan": (C.c_int, [C.c_char_p, _i64, C.POINTER(TsvInfo)]),
    "impop_tsv_fill": (C.c_int, [C.c_char_p, _i64, _p, _p, _p]),
    "impop_gfa_scan": (C.c_int, [C.c_char_p, _i64, C.POINTER(GfaInfo)]),
    "impop_gfa_fill": (C.c_int, [C.c_char_p, _i64, _i32, _p, _p, _p, _p, _p, C.POINTER(_i64), C.POINTER(_i32)]),
    "impop_compact_scan": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _i32, C.c_uint32, _p, _p]),
    "impop_compact_fill": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _i32, C.c_uint32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: {ERRORS.get(code, code)}{(' - ' + detail) if detail else ''}")


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) libimpop_b200.so.  Raises if impossible."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if _build.stale():
            try:
                path = _build.build()
            except Exception as exc:  # no nvcc and no prebuilt library: nothing to fall back to
                if not os.path.exists(path):
                    raise RuntimeError(
                        "libimpop_b200.so is missing and could not be built (nvcc unavailable); "
                        "impop_b200 has no CPU fallback") from exc
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the ABI lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
