"""Device engine: thin Python objects over the C ABI of libimpop_b200.so (include/impop_b200.h).

PyTorch is used only for device memory, streams and (in distributed.py) the process group;
every computation below is a call into the hand-written sm_100a library.  There is no CPU
fallback: constructing a Context without a CUDA device raises.

torch is imported lazily: window batches (matrix mode, bench, multi-GPU) use torch tensors; the TSV-mode
reductions also take plain `devmem.DevArray` buffers, so that the per-window command lines never pay
the import (`Context(lite=True)`).
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _native as N
from .devmem import DevArray


class _LazyTorch:
    def __getattr__(self, name):
        import torch as _t
        globals()["torch"] = _t
        return getattr(_t, name)


torch = _LazyTorch()

_NP = {"f64": np.float64, "i64": np.int64, "i32": np.int32, "u8": np.uint8}

LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = N.LAB_SUBSET, N.LAB_A, N.LAB_B, N.LAB_SEG
ALGO_TCGEN05, ALGO_SIMT = N.ALGO_TCGEN05, N.ALGO_SIMT
NSTATS, NCOUNTS, ST = N.NSTATS, N.NCOUNTS, N.ST


def _ptr(t):
    """Raw pointer of a torch tensor / DevArray / numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _stream_ptr(stream=None):
    if stream is None:
        if "torch" not in sys.modules:                      # lite contexts: the default stream
            return None
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def _is_f64(t) -> bool:
    return t.dtype == np.float64 if isinstance(t, DevArray) else t.dtype == torch.float64


def _u32_tensor(a: np.ndarray) -> torch.Tensor:
    """numpy uint32 -> torch int32 view (torch's uint32 support is partial; the bits are what matter)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32))


def _u64_tensor(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64))


class Context:
    """One impop_ctx_t per device (impop_create / impop_destroy)."""

    def __init__(self, device: int = 0, lite: bool = False):
        """lite=True: no torch import (TSV-mode command lines); arrays are devmem.DevArray buffers."""
        self.lib = N.lib()
        self.device = int(device)
        self.lite = bool(lite)
        if not lite:
            if not torch.cuda.is_available():
                raise RuntimeError("impop_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            torch.cuda.set_device(self.device)
            torch.zeros(1, device=self.torch_device)      # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.impop_create(self.device, C.byref(h))
        if rc != 0:
            raise N.NativeError(rc, "impop_create", "no sm_100a device or CUDA failure")
        self.handle = h

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.impop_destroy(self.handle)
            self.handle = None

    def _empty(self, shape, kind: str, like=None, zero: bool = False):
        """Uninitialised (or zeroed) device array of the caller's flavour: DevArray next to DevArrays / on lite contexts."""
        if self.lite or isinstance(like, DevArray):
            out = DevArray(self, shape, _NP[kind])
            return out.zero_() if zero else out
        dt = {"f64": torch.float64, "i64": torch.int64, "i32": torch.int32, "u8": torch.uint8}[kind]
        dev = like.device if like is not None and hasattr(like, "device") else self.torch_device
        return (torch.zeros if zero else torch.empty)(shape, dtype=dt, device=dev)

    def upload(self, a: np.ndarray):
        """Host array -> device array of this context's flavour."""
        if self.lite:
            return DevArray.from_numpy(self, a)
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.torch_device)

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            raise N.NativeError(rc, name, self.lib.impop_last_error(self.handle).decode())

    def check(self, stream=None):
        """Synchronise the stream and raise on sticky device-side errors (impop_check)."""
        self._call("impop_check", _stream_ptr(stream))

    @property
    def launches(self) -> int:
        return int(self.lib.impop_launch_count(self.handle))

    def timing(self, enable: bool = True):
        """Start (or stop) recording per-kernel CUDA-event timings; clears earlier records."""
        self._call("impop_timing_enable", 1 if enable else 0)

    def timing_read(self, kernel: str):
        """(total milliseconds, launches) of one kernel (see _native.KERNELS) since timing(True)."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        self._call("impop_timing_read", N.KERNELS[kernel], C.byref(ms), C.byref(cnt))
        return ms.value, cnt.value

    # ------------------------------------------------------------------ stand-alone kernels
    def pack_bits(self, dense: torch.Tensor, pitch_words: int | None = None, stream=None) -> torch.Tensor:
        """K1: dense 0/1 uint8 [n, m] (device) -> bit-packed int32 [n, pitch_words]."""
        assert dense.dtype == torch.uint8 and dense.dim() == 2 and dense.is_cuda
        n, m = dense.shape
        if pitch_words is None:
            pitch_words = ((m + 127) // 128) * 4
        out = torch.empty((n, pitch_words), dtype=torch.int32, device=dense.device)
        self._call("impop_pack_bits", _ptr(dense), n, m, dense.stride(0) if n else max(m, 1), _ptr(out), pitch_words,
                   _stream_ptr(stream))
        return out

    def reduce_identity(self, ident: torch.Tensor, labels: torch.Tensor | None, weight: torch.Tensor | None = None,
                        length: int = 0, seg_sites: float = 0.0, stream=None):
        """K3 alone (TSV mode): (stats[NSTATS] f64, counts[NCOUNTS] i64, wsum[4] f64) device tensors.
        wsum = weighted sum, weighted pair count, grouped pi, grouped pi / length (see include/impop_b200.h)."""
        assert _is_f64(ident) and ident.dim() == 2 and ident.is_cuda
        n = ident.shape[0]
        stats = self._empty(NSTATS, "f64", ident)
        counts = self._empty(NCOUNTS, "i64", ident)
        wsum = self._empty(4, "f64", ident, zero=True)
        self._call("impop_reduce_identity", _ptr(ident), n, ident.stride(0) if n else 0, _ptr(labels), _ptr(weight),
                   int(length or 0), float(seg_sites), _ptr(stats), _ptr(counts), _ptr(wsum), _stream_ptr(stream))
        return stats, counts, wsum

    def tajima_d(self, n: torch.Tensor, S: torch.Tensor, pi: torch.Tensor, with_parts: bool = False, stream=None):
        count = n.numel()
        D = self._empty(count, "f64", n)
        parts = self._empty((count, 10), "f64", n) if with_parts else None
        self._call("impop_tajima_d", _ptr(n), _ptr(S), _ptr(pi), count, _ptr(D), _ptr(parts), _stream_ptr(stream))
        return (D, parts) if with_parts else D

    def site_counts(self, sites: torch.Tensor, masks: torch.Tensor, want_freq: bool = True, stream=None,
                    out_counts: torch.Tensor | None = None, out_freq: torch.Tensor | None = None):
        """K4: sites [M, words] int64 (u64 bits), masks [P, words] -> counts [M, P] i32, freq [M, P] f64."""
        M, words = sites.shape
        P = masks.shape[0]
        counts = out_counts if out_counts is not None else self._empty((M, P), "i32", sites)
        freq = out_freq if out_freq is not None else (self._empty((M, P), "f64", sites) if want_freq else None)
        self._call("impop_site_counts", _ptr(sites), M, words, _ptr(masks), P, _ptr(counts), _ptr(freq),
                   _stream_ptr(stream))
        return counts, freq

    def round_decimal(self, values, digits: int, stream=None):
        """CPython's round(x, digits) on every element of a device fp64 array, in place (impop_round_decimal)."""
        self._call("impop_round_decimal", _ptr(values), values.numel(), int(digits), _stream_ptr(stream))
        return values

    def repitch_rows(self, src, dst, rows, src_pitch, dst_pitch, src_off, dst_off, stream=None):
        """Tight rows (as stored / transferred: src_pitch words per haplotype) -> the 16-byte-multiple rows the kernels read,
        zero padded (impop_repitch_rows).  src / dst: device u32 (int32) arrays; the tables are host arrays, one entry per window."""
        tabs = [np.ascontiguousarray(a, dtype=dt) for a, dt in ((rows, np.int32), (src_pitch, np.int32), (dst_pitch, np.int32),
                                                                (src_off, np.int64), (dst_off, np.int64))]
        self._call("impop_repitch_rows", int(tabs[0].shape[0]), *[_ptr(t) for t in tabs], _ptr(src), _ptr(dst), _stream_ptr(stream))
        return dst

    def selftest_division(self, count: int = 1 << 24, seed: int = 1, stream=None) -> int:
        """Mismatches between the epilogue's in-range division and __ddiv_rn over `count` random triples."""
        bad = C.c_int64(-1)
        self._call("impop_selftest_division", C.c_uint64(seed), int(count), C.byref(bad), _stream_ptr(stream))
        return int(bad.value)

    def greedy_groups(self, ident: torch.Tensor, threshold: float, stream=None):
        """pica2 step 1 on the device: (group [n] i32 = seed index, weight [n] f64 = |G|/n on seeds)."""
        n = ident.shape[0]
        group = self._empty(n, "i32", ident)
        weight = self._empty(n, "f64", ident)
        self._call("impop_greedy_groups", _ptr(ident), n, ident.stride(0) if n else 0, float(threshold), _ptr(group),
                   _ptr(weight), _stream_ptr(stream))
        return group, weight

    def cluster(self, ident: torch.Tensor, threshold: float, stream=None) -> torch.Tensor:
        """K5: component label (= smallest member index) per row of a dense identity matrix."""
        n = ident.shape[0]
        comp = self._empty(n, "i32", ident)
        self._call("impop_cluster", _ptr(ident), n, ident.stride(0) if n else 0, float(threshold), _ptr(comp),
                   _stream_ptr(stream))
        return comp


class WindowBatch:
    """A batch of windows resident on the device (impop_batch_t).

    x: int32 tensor holding the u32 presence words of every window; node_len: int32 (u32) tensor;
    labels: uint8 tensor.  Per-window descriptor arrays are host numpy arrays.
    """

    def __init__(self, ctx: Context, n, m, pitch_words, x_off, len_off, lab_off, length,
                 x: torch.Tensor, node_len: torch.Tensor, labels: torch.Tensor, node_len_host=None, stream=None, site_runs=None,
                 row_adj=None, win_const=None, col_mult=None, heavy_entries=None):
        self.ctx = ctx
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.m = np.ascontiguousarray(m, dtype=np.int32)
        self.pitch_words = np.ascontiguousarray(pitch_words, dtype=np.int32)
        self.x_off = np.ascontiguousarray(x_off, dtype=np.int64)
        self.len_off = np.ascontiguousarray(len_off, dtype=np.int64)
        self.lab_off = np.ascontiguousarray(lab_off, dtype=np.int64)
        self.length = np.ascontiguousarray(length, dtype=np.int64)
        self.windows = int(self.n.shape[0])
        self.x, self.node_len, self.labels = x, node_len, labels      # keep the device buffers alive
        # site_runs: IMPOP_ST_S_BUBBLES per window as counted at ingest on the original node order (None: counted on the device)
        self.site_runs = None if site_runs is None else np.ascontiguousarray(site_runs, dtype=np.int64)
        # affine form of the windows (ingest.compact_batch(pairs=True)): row terms R_i (device int32, the windows' rows in batch order), window
        # constants C (host int64), column multiplicities for S (device uint8, columns at len_off); None: a plain batch
        self.row_adj, self.col_mult = row_adj, col_mult
        self.win_const = None if win_const is None else np.ascontiguousarray(win_const, dtype=np.int64)
        # heavy_entries: per window, sum over its nodes of ceil(floor(len / 255) / 255), when known from ingest (heavy_entries()
        # below): batch set-up then skips its host pass over the node lengths
        self.heavy_entries = None if heavy_entries is None else np.ascontiguousarray(heavy_entries, dtype=np.int32)
        d = N.BatchDesc(self.windows, _ptr(self.n), _ptr(self.m), _ptr(self.pitch_words), _ptr(self.x_off),
                        _ptr(self.len_off), _ptr(self.lab_off), _ptr(self.length), _ptr(x), _ptr(node_len),
                        _ptr(labels), _ptr(node_len_host), _stream_ptr(stream), _ptr(self.site_runs),
                        _ptr(row_adj), _ptr(self.win_const), _ptr(col_mult), _ptr(self.heavy_entries))
        h = C.c_void_p()
        rc = ctx.lib.impop_batch_create(ctx.handle, C.byref(d), C.byref(h))
        if rc != 0:
            raise N.NativeError(rc, "impop_batch_create", ctx.lib.impop_last_error(ctx.handle).decode())
        self.handle = h

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_uniform(cls, ctx: Context, x_bits, node_len, labels, length, m: int | None = None, node_len_host=None,
                     stream=None, site_runs=None, row_adj=None, win_const=None, col_mult=None, heavy_entries=None):
        """W same-shape windows: x_bits [W, n, pitch] u32, node_len [W, m_pad] u32, labels [n] or [W, n] u8; m: columns in use,
        one number or one per window (default: all m_pad; columns beyond m must have length 0 and cost nothing when given).

        Arguments may be numpy arrays (copied to the device) or device tensors (used in place).  Affine form
        (ingest.compact_uniform(pairs=True)): row_adj [W, n] i32, win_const [W] i64 (host), col_mult [W, m_pad] u8.
        """
        W, n, pitch = x_bits.shape
        m_pad = node_len.shape[1]
        as_dev = lambda a, dt: ctx.upload(np.ascontiguousarray(a, dtype=dt).view(np.int32 if dt == np.uint32 else dt)) if isinstance(a, np.ndarray) else a
        xd, ld, lab = as_dev(x_bits, np.uint32), as_dev(node_len, np.uint32), as_dev(labels, np.uint8)
        per_window_labels = lab.dim() == 2
        if row_adj is not None:
            row_adj = as_dev(row_adj, np.int32)
        if col_mult is not None:
            col_mult = as_dev(col_mult, np.uint8)
        ar = np.arange(W, dtype=np.int64)
        L = np.full(W, int(length or 0), dtype=np.int64) if np.isscalar(length) or length is None else np.asarray(length)
        m_w = np.full(W, m_pad) if m is None else (np.full(W, int(m)) if np.isscalar(m) else np.asarray(m, dtype=np.int32))
        return cls(ctx, np.full(W, n), m_w, np.full(W, pitch), ar * (n * pitch),
                   ar * m_pad, ar * n if per_window_labels else np.zeros(W, dtype=np.int64), L, xd, ld, lab,
                   node_len_host=node_len_host, stream=stream, site_runs=site_runs, row_adj=row_adj, win_const=win_const,
                   col_mult=col_mult, heavy_entries=heavy_entries)

    @classmethod
    def from_windows(cls, ctx: Context, windows, site_runs=None):
        """Ragged batch from a list of (x_bits [n, pitch] u32, node_len [m] u32, labels [n] u8, L) host arrays; a window
        in affine form (ingest.compact_window) appends (row_adj [n] i32, win_const, col_mult [m] u8)."""
        n, m, pitch, x_off, len_off, lab_off, L = [], [], [], [], [], [], []
        xs, ls, labs = [], [], []
        radj, wconst, cmult = [], [], []
        affine = any(len(wd) > 4 and wd[4] is not None for wd in windows)
        xo = lo = bo = 0
        for wd in windows:
            xb, nl, lab, length = wd[:4]
            if affine:
                has = len(wd) > 4 and wd[4] is not None
                radj.append(np.asarray(wd[4], dtype=np.int32) if has else np.zeros(len(lab), dtype=np.int32))
                wconst.append(int(wd[5]) if has else 0)
                cmult.append(np.asarray(wd[6], dtype=np.uint8) if has else np.ones(len(nl), dtype=np.uint8))
            xb = np.ascontiguousarray(xb, dtype=np.uint32)
            if xb.ndim != 2:
                xb = xb.reshape(len(lab), -1)
            nn, pw = xb.shape
            if pw % 4:                                     # rows must be 16-byte multiples
                pad = np.zeros((nn, (pw + 3) // 4 * 4), dtype=np.uint32)
                pad[:, :pw] = xb
                xb, pw = pad, pad.shape[1]
            n.append(nn); m.append(len(nl)); pitch.append(pw)
            x_off.append(xo); len_off.append(lo); lab_off.append(bo); L.append(int(length or 0))
            xs.append(xb.ravel()); ls.append(np.asarray(nl, dtype=np.uint32)); labs.append(np.asarray(lab, dtype=np.uint8))
            xo += xb.size
            lo += len(nl)
            bo += nn
        cat = lambda parts, dt: np.concatenate(parts).astype(dt, copy=False) if parts and sum(p.size for p in parts) else np.zeros(4, dtype=dt)
        x = ctx.upload(cat(xs, np.uint32).view(np.int32))
        nl = ctx.upload(cat(ls, np.uint32).view(np.int32))
        lab = ctx.upload(cat(labs, np.uint8))
        extra = {}
        if affine:
            extra = dict(row_adj=ctx.upload(cat(radj, np.int32)), win_const=np.array(wconst, dtype=np.int64),
                         col_mult=ctx.upload(cat(cmult, np.uint8)))
        # heavy-table sizes from the host copies at hand: batch set-up then needs neither a device pass nor a host pass of its own
        heavy = np.array([int(((l.astype(np.int64) // 255 + 254) // 255).sum()) for l in ls], dtype=np.int32)
        return cls(ctx, n, m, pitch, x_off, len_off, lab_off, L, x, nl, lab, site_runs=site_runs, heavy_entries=heavy, **extra)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.impop_batch_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def items(self) -> int:
        return int(self.ctx.lib.impop_batch_items(self.handle))

    # ------------------------------------------------------------------ kernels
    def stats(self, algo: int = ALGO_TCGEN05, stream=None, out_stats=None, out_counts=None):
        """Fused K2+K3 over every window: (stats [W, NSTATS] f64, counts [W, NCOUNTS] i64) on the device."""
        stats = out_stats if out_stats is not None else self.ctx._empty((self.windows, NSTATS), "f64")
        counts = out_counts if out_counts is not None else self.ctx._empty((self.windows, NCOUNTS), "i64")
        self.ctx._call("impop_window_stats", self.handle, algo, _ptr(stats), _ptr(counts), _stream_ptr(stream))
        return stats, counts

    def window_sums(self, rank: int, world: int, algo: int = ALGO_TCGEN05, stream=None) -> torch.Tensor:
        """Raw pair sums [W, 4] over the work items t with t % world == rank (tile-grid split)."""
        sums = self.ctx._empty((self.windows, 4), "f64")
        self.ctx._call("impop_window_sums", self.handle, algo, rank, world, _ptr(sums), _stream_ptr(stream))
        return sums

    def finalize(self, sums_parts: torch.Tensor, stream=None):
        """sums_parts [parts, W, 4] -> (stats, counts); parts are added in index order (reproducible)."""
        assert sums_parts.dim() == 3 and tuple(sums_parts.shape[1:]) == (self.windows, 4)
        stats = self.ctx._empty((self.windows, NSTATS), "f64")
        counts = self.ctx._empty((self.windows, NCOUNTS), "i64")
        self.ctx._call("impop_window_finalize", self.handle, _ptr(sums_parts), sums_parts.shape[0], _ptr(stats),
                       _ptr(counts), _stream_ptr(stream))
        return stats, counts

    def pairwise(self, window: int, algo: int = ALGO_TCGEN05, want_i=True, want_pi=True, stream=None):
        """Materialise one window: (I [n, n] i64, A [n] i64, pi [n, n] f64); the table `impg similarity` prints."""
        n = int(self.n[window])
        I = self.ctx._empty((n, n), "i64", zero=True) if want_i else None
        A = self.ctx._empty(n, "i64", zero=True)
        pi = self.ctx._empty((n, n), "f64", zero=True) if want_pi else None
        self.ctx._call("impop_pairwise", self.handle, window, algo, _ptr(I), _ptr(A), _ptr(pi), _stream_ptr(stream))
        return I, A, pi
