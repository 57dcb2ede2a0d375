"""ctypes loader for oracle/liboracle_impop.so (plain-C oracle).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_impop.so")
NSTATS, NCOUNTS = 20, 8
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "csrc", "oracle_impop.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle_impop.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        p = C.c_void_p
        L.oracle_window_pairwise.argtypes = [p, C.c_int, C.c_int, C.c_int, p, p, p, p, C.c_int]
        L.oracle_window_stats.argtypes = [p, C.c_int, C.c_int, C.c_int, p, p, C.c_int64, p, p]
        L.oracle_batch_stats.argtypes = [C.c_int, p, p, p, p, p, p, p, p, p, p, p, p, C.c_int]
        L.oracle_tajimas_d.argtypes = [C.c_int64, C.c_double, C.c_double, p]
        L.oracle_tajimas_d.restype = C.c_double
        L.oracle_pi_from_counts.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.oracle_pi_from_counts.restype = C.c_double
        L.oracle_site_counts.argtypes = [p, C.c_int64, C.c_int, p, C.c_int, p, p]
        L.oracle_site_counts.restype = None
        L.oracle_finalize.argtypes = [p, p, C.c_int64, p]
        L.oracle_finalize.restype = None
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def window_pairwise(x_bits: np.ndarray, m: int, node_len: np.ndarray, want_i=True, want_pi=True, use_lut=True):
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint32)
    n, pitch = x_bits.shape
    node_len = np.ascontiguousarray(node_len, dtype=np.uint32)
    A = np.zeros(n, dtype=np.int64)
    I = np.zeros((n, n), dtype=np.int64) if want_i else None
    pi = np.zeros((n, n), dtype=np.float64) if want_pi else None
    rc = lib().oracle_window_pairwise(_ptr(x_bits), n, m, pitch, _ptr(node_len), _ptr(A), _ptr(I), _ptr(pi), int(use_lut))
    assert rc == 0
    return A, I, pi


def window_stats(x_bits: np.ndarray, m: int, node_len: np.ndarray, labels: np.ndarray, L: int):
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint32)
    n, pitch = x_bits.shape
    node_len = np.ascontiguousarray(node_len, dtype=np.uint32)
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    stats = np.zeros(NSTATS, dtype=np.float64)
    counts = np.zeros(NCOUNTS, dtype=np.int64)
    rc = lib().oracle_window_stats(_ptr(x_bits), n, m, pitch, _ptr(node_len), _ptr(labels), int(L or 0), _ptr(stats), _ptr(counts))
    assert rc == 0
    return stats, counts


def batch_stats(n, m, pitch, x_off, len_off, lab_off, L, x, node_len, labels, threads: int):
    W = len(n)
    arrs = [np.ascontiguousarray(a, dtype=t) for a, t in
            ((n, np.int32), (m, np.int32), (pitch, np.int32), (x_off, np.int64), (len_off, np.int64),
             (lab_off, np.int64), (L, np.int64), (x, np.uint32), (node_len, np.uint32), (labels, np.uint8))]
    stats = np.zeros((W, NSTATS), dtype=np.float64)
    counts = np.zeros((W, NCOUNTS), dtype=np.int64)
    rc = lib().oracle_batch_stats(W, *[_ptr(a) for a in arrs], _ptr(stats), _ptr(counts), int(threads))
    assert rc == 0
    return stats, counts


def tajimas_d(n: int, S: float, pi: float):
    parts = np.zeros(10, dtype=np.float64)
    d = lib().oracle_tajimas_d(int(n), float(S), float(pi), _ptr(parts))
    return d, parts


def site_counts(sites: np.ndarray, masks: np.ndarray):
    sites = np.ascontiguousarray(sites, dtype=np.uint64)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    M, words = sites.shape
    P = masks.shape[0]
    counts = np.zeros((M, P), dtype=np.int32)
    freq = np.zeros((M, P), dtype=np.float64)
    lib().oracle_site_counts(_ptr(sites), M, words, _ptr(masks), P, _ptr(counts), _ptr(freq))
    return counts, freq
