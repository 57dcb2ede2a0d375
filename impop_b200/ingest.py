"""Ingest for matrix mode: window graphs (GFA v1) -> bit-packed haplotype x node matrices + node lengths + names.

The reference never holds a matrix: per window it extracts a graph and hands *text* to the similarity tool
(`odgi similarity -i tmp.gfa`, run_pica2_odgi.sh:60-96; `impg similarity -r REGION`, run_h-fst.sh:65-67), whose
all-pairs TSV the scripts parse again.  Here the window graph is parsed once by libimpop_b200's host-side reader
(`impop_gfa_scan` / `impop_gfa_fill`, include/impop_b200.h) into exactly the arrays `WindowBatch.from_windows`
uploads, and a set of windows can be kept in one compact binary container so that nothing round-trips through
text between extraction and the GPU.  Only parsing and name handling happen on the host.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native
from ._native import GfaInfo, NativeError, lib


@dataclass
class GraphWindow:
    """One window: rows = paths (file order), columns = segments (file order)."""
    names: list            # row names, e.g. 'HG00097#1#CM094061.1:109468899-109469099'
    x_bits: np.ndarray     # [n, pitch_words] uint32, bit k & 31 of word k >> 5 = node k present
    node_len: np.ndarray   # [m] uint32
    counts: np.ndarray | None = None   # [n, m] uint16 visit counts (multiset coverage), when requested
    region: str | None = None
    length: int = 0        # window length L in bp (BED end - start); 0 = unknown
    site_runs: int = -1    # variant sites (bubble-like runs of segregating nodes) counted before compaction; -1: not counted
    row_adj: np.ndarray | None = None   # affine form (compact_window): R_i [n] int32, window constant C, column multiplicities [m] uint8
    win_const: int = 0
    col_mult: np.ndarray | None = None
    revisits: bool = False  # some path visits a node more than once (reported by the reader; multiset coverage then differs from presence)

    @property
    def n(self) -> int:
        return self.x_bits.shape[0]

    @property
    def m(self) -> int:
        return self.node_len.shape[0]


def _pitch_for(m: int) -> int:
    return max(4, ((m + 127) // 128) * 4)


def parse_gfa(text, want_counts=False, region: str | None = None, length: int = 0) -> GraphWindow:
    """GFA v1 text (bytes / str / path-like object with .read) -> GraphWindow.  Raises NativeError with the
    offending line number on malformed input.  want_counts: True / False, or "auto" = visit counts only when some path
    revisits a node (the window is then read a second time: rare, and a chromosome's worth of count matrices is
    gigabytes)."""
    if hasattr(text, "read"):
        text = text.read()
    if isinstance(text, str):
        text = text.encode()
    L = lib()
    info = GfaInfo()
    rc = L.impop_gfa_scan(text, len(text), C.byref(info))
    if rc:
        raise NativeError(rc, "impop_gfa_scan", f"malformed GFA at line {info.error_line}")
    n, m = int(info.paths), int(info.segments)
    pitch = _pitch_for(m)
    x = np.zeros((n, pitch), dtype=np.uint32)
    node_len = np.zeros(m, dtype=np.uint32)
    counts = np.zeros((n, m), dtype=np.uint16) if want_counts is True else None
    names_buf = np.zeros(max(int(info.name_bytes), 1), dtype=np.uint8)
    name_off = np.zeros(n + 1, dtype=np.int64)
    err_line, revisits = C.c_int64(0), C.c_int32(0)
    rc = L.impop_gfa_fill(text, len(text), pitch, x.ctypes.data, node_len.ctypes.data,
                          counts.ctypes.data if counts is not None else None, names_buf.ctypes.data,
                          name_off.ctypes.data, C.byref(err_line), C.byref(revisits))
    if rc:
        raise NativeError(rc, "impop_gfa_fill", f"malformed GFA at line {err_line.value}")
    if want_counts == "auto" and revisits.value:
        return parse_gfa(text, True, region, length)
    raw = names_buf.tobytes()
    names = [raw[name_off[i]:name_off[i + 1] - 1].decode() for i in range(n)]
    win = GraphWindow(names, x, node_len, counts, region, int(length))
    win.revisits = bool(revisits.value)
    return win


def read_gfa(path, want_counts=False, region: str | None = None, length: int = 0) -> GraphWindow:
    with open(path, "rb") as fh:
        return parse_gfa(fh.read(), want_counts, region, length)


def read_gfa_many(items, threads: int | None = None, want_counts=False) -> list:
    """[(region, path, length), ...] -> [GraphWindow, ...] in the same order.  The native reader runs outside the GIL
    (ctypes), so the window graphs of a chromosome are parsed on all host cores; `threads=1` reads one by one."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    items = list(items)
    if threads is None:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(int(threads), len(items) or 1))
    one = lambda it: read_gfa(it[1], want_counts=want_counts, region=it[0], length=int(it[2] or 0))
    if threads == 1:
        return [one(it) for it in items]
    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(one, items))                 # map keeps the input order; the first exception propagates


def write_gfa(handle, names, x: np.ndarray, node_len: np.ndarray, walks: bool = False) -> None:
    """Emit a window as GFA v1 (what the synthetic generator hands to the text path and to tests): one S line per
    node (sequence 'N' * len up to 64 bp, else '*' + LN:i:), one P (or W) line per haplotype listing its nodes."""
    x = np.asarray(x).astype(bool)
    handle.write("H\tVN:Z:1.0\n")
    for k, ln in enumerate(np.asarray(node_len).tolist()):
        if 0 < ln <= 64:
            handle.write(f"S\t{k + 1}\t{'N' * ln}\n")
        else:
            handle.write(f"S\t{k + 1}\t*\tLN:i:{ln}\n")
    for name, row in zip(names, x):
        idx = np.flatnonzero(row) + 1
        if walks:
            sample, hap, rest = name.split("#", 2)
            seqid, _, rng = rest.partition(":")
            s, _, e = rng.partition("-")
            handle.write(f"W\t{sample}\t{hap}\t{seqid}\t{s or '*'}\t{e or '*'}\t" + "".join(f">{v}" for v in idx) + "\n")
        else:
            handle.write(f"P\t{name}\t" + (",".join(f"{v}+" for v in idx) if len(idx) else "*") + "\t*\n")


# ------------------------------------------------------------------------------------------------------------
# Compact container for a set of windows (one .npz: every array concatenated, offsets per window)
# ------------------------------------------------------------------------------------------------------------
def save_batch(path, windows) -> None:
    """Write windows (GraphWindow list) as one binary file: bit matrices, node lengths, names, regions, lengths."""
    n = np.array([w.n for w in windows], dtype=np.int32)
    m = np.array([w.m for w in windows], dtype=np.int32)
    pitch = np.array([w.x_bits.shape[1] for w in windows], dtype=np.int32)
    x = np.concatenate([w.x_bits.reshape(-1) for w in windows]) if windows else np.zeros(0, np.uint32)
    nl = np.concatenate([w.node_len for w in windows]) if windows else np.zeros(0, np.uint32)
    names = "\n".join("\t".join(w.names) for w in windows)
    regions = "\n".join(w.region or "" for w in windows)
    L = np.array([w.length for w in windows], dtype=np.int64)
    radj, wconst, cmult, runs = _affine_arrays(windows)
    np.savez(path, format=np.array([2]), n=n, m=m, pitch=pitch, x=x.astype(np.uint32), node_len=nl.astype(np.uint32),
             names=np.frombuffer(names.encode(), dtype=np.uint8), regions=np.frombuffer(regions.encode(), dtype=np.uint8),
             length=L, row_adj=radj, win_const=wconst, col_mult=cmult, site_runs=runs)


def load_batch(path) -> list:
    z = np.load(path)
    n, m, pitch = z["n"], z["m"], z["pitch"]
    x_all, len_all, L_all = z["x"], z["node_len"], z["length"]        # NpzFile re-reads a member on every access: once each
    names = z["names"].tobytes().decode().split("\n") if len(n) else []
    regions = z["regions"].tobytes().decode().split("\n") if len(n) else []
    affine = "row_adj" in z.files
    if affine:
        radj, wconst, cmult, runs = z["row_adj"], z["win_const"], z["col_mult"], z["site_runs"]
    out, xo, lo, ro = [], 0, 0, 0
    for w in range(len(n)):
        xs = int(n[w]) * int(pitch[w])
        g = GraphWindow(names[w].split("\t") if names[w] else [],
                        x_all[xo:xo + xs].reshape(int(n[w]), int(pitch[w])).copy(),
                        len_all[lo:lo + int(m[w])].copy(), None, regions[w] or None, int(L_all[w]))
        if affine:
            g.site_runs = int(runs[w])
            g.row_adj, g.win_const, g.col_mult = radj[ro:ro + int(n[w])].copy(), int(wconst[w]), cmult[lo:lo + int(m[w])].copy()
        out.append(g)
        xo += xs
        lo += int(m[w])
        ro += int(n[w])
    return out


def _affine_arrays(windows):
    """(row_adj int32 [rows], win_const int64 [W], col_mult uint8 [nodes], site_runs int64 [W]) of a GraphWindow list; a
    plain window contributes R = 0, C = 0, multiplicity 1."""
    radj = [np.asarray(w.row_adj, dtype=np.int32) if w.row_adj is not None else np.zeros(w.n, dtype=np.int32) for w in windows]
    cmult = [np.asarray(w.col_mult, dtype=np.uint8) if w.col_mult is not None else np.ones(w.m, dtype=np.uint8) for w in windows]
    cat = lambda parts, dt: np.concatenate(parts).astype(dt, copy=False) if parts else np.zeros(0, dtype=dt)
    return (cat(radj, np.int32), np.array([int(w.win_const) for w in windows], dtype=np.int64), cat(cmult, np.uint8),
            np.array([int(w.site_runs) for w in windows], dtype=np.int64))


# ------------------------------------------------------------------------------------------------------------
# Flat container: one file, fixed little-endian layout, every array usable in place (np.memmap) -- a chromosome's
# windows go from disk to one WindowBatch without a per-window Python loop and without copies.
#   header  8 x int64: magic, version (2), W, rows (sum n), x words, nodes (sum m), unique names U, text bytes
#   int64   n[W] m[W] pitch[W] x_off[W] len_off[W] row_off[W] length[W] site_runs[W] win_const[W]
#   int32   name_id[rows]            index of the row's haplotype in the unique-name table
#   int32   row_adj[rows]            affine form (impop_batch_desc_t): R_i; a plain window has R = 0, C = 0, multiplicity 1
#   uint32  node_len[nodes]
#   uint8   col_mult[nodes]          (padded to 4 bytes)
#   uint32  x[x words]               (starts on a 64-byte boundary)
#   text    unique names ('\n' joined; a row name = unique name + ':' + the window's coordinates when the source had them)
#           + '\x00' + regions ('\n' joined) + '\x00' + per-window coordinate suffixes ('\n' joined)
# ------------------------------------------------------------------------------------------------------------
FLAT_MAGIC = 0x31574F504D49          # "IMPOW1"


@dataclass
class FlatBatch:
    n: np.ndarray
    m: np.ndarray
    pitch: np.ndarray
    x_off: np.ndarray
    len_off: np.ndarray
    row_off: np.ndarray
    length: np.ndarray
    name_id: np.ndarray
    node_len: np.ndarray
    x: np.ndarray
    uniq: list
    regions: list
    suffix: list
    site_runs: np.ndarray | None = None
    win_const: np.ndarray | None = None
    row_adj: np.ndarray | None = None
    col_mult: np.ndarray | None = None

    @property
    def windows(self) -> int:
        return int(self.n.shape[0])

    def names(self, w: int) -> list:
        ids = self.name_id[int(self.row_off[w]):int(self.row_off[w]) + int(self.n[w])]
        sfx = self.suffix[w]
        return [self.uniq[i] + ((":" + sfx) if sfx else "") for i in ids]

    def labels(self, pop_a=None, pop_b=None, subset=None, seg=None) -> np.ndarray:
        """Label byte per row of the whole batch from PREFIX lists (after h-fst.py:18-61 canonicalisation): the classes
        are decided once per unique haplotype name and gathered per row."""
        def flags(prefixes, bit, default):
            if prefixes is None:
                return np.full(len(self.uniq), bit if default else 0, dtype=np.uint8)
            pre = tuple(prefixes)
            return np.array([bit if u.startswith(pre) else 0 for u in self.uniq], dtype=np.uint8) if pre else np.zeros(len(self.uniq), np.uint8)
        per = (flags(subset, _native.LAB_SUBSET, True) | flags(seg if seg is not None else subset, _native.LAB_SEG, True)
               | flags(pop_a, _native.LAB_A, False) | flags(pop_b, _native.LAB_B, False))
        return per[self.name_id]

    def window(self, w: int) -> GraphWindow:
        xs, pw, nn = int(self.x_off[w]), int(self.pitch[w]), int(self.n[w])
        lo = int(self.len_off[w])
        g = GraphWindow(self.names(w), np.asarray(self.x[xs:xs + nn * pw]).reshape(nn, pw), np.asarray(self.node_len[lo:lo + int(self.m[w])]),
                        None, self.regions[w] or None, int(self.length[w]))
        if self.row_adj is not None:
            ro = int(self.row_off[w])
            g.site_runs = int(self.site_runs[w])
            g.row_adj, g.win_const = np.asarray(self.row_adj[ro:ro + nn]), int(self.win_const[w])
            g.col_mult = np.asarray(self.col_mult[lo:lo + int(self.m[w])])
        return g


def _split_name(name: str):
    head, sep, tail = name.rpartition(":")
    if sep and "-" in tail and tail.replace("-", "").isdigit():
        return head, tail
    return name, ""


def save_flat(path, windows) -> None:
    """GraphWindow list -> flat container.  Row names are stored once per haplotype (the part before ':start-end')."""
    W = len(windows)
    n = np.array([w.n for w in windows], dtype=np.int64)
    m = np.array([w.m for w in windows], dtype=np.int64)
    pitch = np.array([w.x_bits.shape[1] for w in windows], dtype=np.int64)
    rows = n * pitch
    x_off = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
    len_off = np.concatenate([[0], np.cumsum(m)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
    row_off = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
    length = np.array([w.length for w in windows], dtype=np.int64)
    uniq, index, ids, suffix = [], {}, [], []
    for w in windows:
        sfx = None
        for name in w.names:
            head, tail = _split_name(name)
            if sfx is None:
                sfx = tail
            elif tail != sfx:                      # mixed coordinates inside one window: keep full names
                head, tail = name, ""
            k = index.get(head)
            if k is None:
                k = index[head] = len(uniq)
                uniq.append(head)
            ids.append(k)
        suffix.append(sfx or "")
    name_id = np.array(ids, dtype=np.int32)
    text = ("\n".join(uniq) + "\x00" + "\n".join(w.region or "" for w in windows) + "\x00" + "\n".join(suffix)).encode()
    node_len = np.concatenate([w.node_len for w in windows]).astype(np.uint32) if W else np.zeros(0, np.uint32)
    x = np.concatenate([w.x_bits.reshape(-1) for w in windows]).astype(np.uint32) if W else np.zeros(0, np.uint32)
    radj, wconst, cmult, runs = _affine_arrays(windows)
    hdr = np.array([FLAT_MAGIC, 2, W, int(n.sum()), int(x.size), int(m.sum()), len(uniq), len(text)], dtype=np.int64)
    with open(path, "wb") as fh:
        fh.write(hdr.tobytes())
        for a in (n, m, pitch, x_off, len_off, row_off, length, runs, wconst):
            fh.write(a.tobytes())
        fh.write(name_id.tobytes())
        fh.write(radj.tobytes())
        fh.write(node_len.tobytes())
        fh.write(cmult.tobytes())
        fh.write(b"\x00" * ((-fh.tell()) % 64))
        fh.write(x.tobytes())
        fh.write(text)


def load_flat(path, mmap: bool = True) -> FlatBatch:
    """Flat container -> FlatBatch whose arrays are views of the file (np.memmap) -- no per-window work."""
    buf = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    hdr = np.frombuffer(buf, dtype=np.int64, count=8)
    if int(hdr[0]) != FLAT_MAGIC or int(hdr[1]) not in (1, 2):
        raise ValueError(f"{path}: not an impop window container")
    v2 = int(hdr[1]) == 2
    W, rows, xw, nodes, U, tb = (int(v) for v in hdr[2:8])
    off = 64
    out = []
    for _ in range(9 if v2 else 7):
        out.append(np.frombuffer(buf, dtype=np.int64, count=W, offset=off))
        off += 8 * W
    runs, wconst = (out.pop(7), out.pop(7)) if v2 else (None, None)
    name_id = np.frombuffer(buf, dtype=np.int32, count=rows, offset=off)
    off += 4 * rows
    radj = cmult = None
    if v2:
        radj = np.frombuffer(buf, dtype=np.int32, count=rows, offset=off)
        off += 4 * rows
    node_len = np.frombuffer(buf, dtype=np.uint32, count=nodes, offset=off)
    off += 4 * nodes
    if v2:
        cmult = np.frombuffer(buf, dtype=np.uint8, count=nodes, offset=off)
        off += nodes
    off += (-off) % 64
    x = np.frombuffer(buf, dtype=np.uint32, count=xw, offset=off)
    off += 4 * xw
    text = bytes(buf[off:off + tb]).decode()
    uniq, regions, suffix = (part.split("\n") if part else [] for part in (text.split("\x00") + ["", ""])[:3])
    if W and not regions:
        regions = [""] * W
    if W and not suffix:
        suffix = [""] * W
    return FlatBatch(*out, name_id, node_len, x, uniq, regions, suffix, runs, wconst, radj, cmult)


def labels_from_names(names, pop_a=None, pop_b=None, subset=None, seg=None) -> np.ndarray:
    """Label byte per row from name sets (after h-fst.py:64-82 expansion / run_tajd.sh -l subsetting):
    SUBSET / SEG default to every row."""
    lab = np.zeros(len(names), dtype=np.uint8)
    for i, s in enumerate(names):
        f = 0
        if subset is None or s in subset:
            f |= _native.LAB_SUBSET
        if seg is None or s in seg:
            f |= _native.LAB_SEG
        if pop_a is not None and s in pop_a:
            f |= _native.LAB_A
        if pop_b is not None and s in pop_b:
            f |= _native.LAB_B
        lab[i] = f
    return lab


def multiset_expand(win: GraphWindow, max_copies: int = 255) -> GraphWindow:
    """Multiset (visit-count) coverage as a presence matrix the set kernels take unchanged.

    With c_ik = how often path i visits node k, the multiset intersection is sum_k len_k min(c_ik, c_jk)
    [UPSTREAM-UNVERIFIED: what `odgi similarity` computes on graphs whose paths revisit nodes].  Because
    min(a, b) = sum_{t >= 1} [a >= t][b >= t], it equals the SET intersection over "copy" nodes (k, t), t = 1 ..
    max_i c_ik, each of length len_k and present in path i iff c_ik >= t.  Path lengths become sum_k len_k c_ik and
    the union A_i + A_j - I_ij follows; every kernel and statistic downstream is untouched.  Needs `win.counts`
    (parse_gfa(..., want_counts=True)); copies beyond `max_copies` are dropped (the reader saturates at 65 535)."""
    if win.counts is None:
        raise ValueError("multiset_expand needs visit counts: parse the GFA with want_counts=True")
    c = np.minimum(win.counts.astype(np.int64), max_copies)
    tmax = c.max(axis=0) if c.size else np.zeros(win.m, dtype=np.int64)
    tmax = np.maximum(tmax, 1)                                   # a node nobody visits stays as one empty column
    node = np.repeat(np.arange(win.m), tmax)                     # copy column -> node
    first = np.cumsum(tmax) - tmax
    t = np.arange(node.shape[0]) - first[node] + 1               # copy index 1 .. tmax
    dense = (c[:, node] >= t[None, :]).astype(np.uint8)
    m2 = node.shape[0]
    pitch = _pitch_for(m2)
    padded = np.zeros((win.n, pitch * 32), dtype=np.uint8)
    padded[:, :m2] = dense
    bits = np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(win.n, pitch).copy()
    return GraphWindow(list(win.names), bits, win.node_len[node].astype(np.uint32), None, win.region, win.length)


# ------------------------------------------------------------------------------------------------------------
# Column compaction (impop_compact_scan / impop_compact_fill): once per window, before the upload
# ------------------------------------------------------------------------------------------------------------
def _host_threads(threads):
    import os
    if threads is None:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, int(threads))


@dataclass
class Compacted:
    """What impop_compact_scan / _fill return for a host batch; the affine form (pairs=True) adds row_adj / win_const /
    col_mult (see impop_batch_desc_t), which a batch built from x_out / len_out MUST be given."""
    m_out: np.ndarray
    pitch_out: np.ndarray
    x_off_out: np.ndarray
    len_off_out: np.ndarray
    x_out: np.ndarray
    len_out: np.ndarray
    site_runs: np.ndarray
    row_off_out: np.ndarray | None = None
    row_adj: np.ndarray | None = None      # int32, the windows' rows in batch order
    win_const: np.ndarray | None = None    # int64 [W]
    col_mult: np.ndarray | None = None     # uint8, laid out like len_out


def compact_batch(n, m, pitch_words, x_off, len_off, x_bits, node_len, threads=None, uniform_pitch: bool = False,
                  pairs: bool = True, replicate: bool = True) -> Compacted:
    """Compact every window of a host batch (descriptor arrays as WindowBatch takes them; x_bits / node_len flat uint32).

    Nodes carried by every haplotype of a window are merged into one node of their summed length, nodes carried by
    none (or of length 0) are dropped, the rest is ordered by length: every I_ij, A_i, U_ij, S and statistic is
    unchanged, while an HPRC-shaped window loses the third of its columns that is backbone.
    pairs=True (affine form): identical columns are merged and the two complementary columns of a bi-allelic bubble
    become one (row terms R_i, window constant C, column multiplicities for S); replicate=True also spreads weights
    >= 255 over copies of their column where the copies fit into the chunk padding (include/impop_b200.h).
    len_out rows are padded with zeros to 32 * pitch_out nodes so that a window's lengths can be sliced like its
    presence words."""
    L = lib()
    n = np.ascontiguousarray(n, dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int32)
    pitch_words = np.ascontiguousarray(pitch_words, dtype=np.int32)
    x_off = np.ascontiguousarray(x_off, dtype=np.int64)
    len_off = np.ascontiguousarray(len_off, dtype=np.int64)
    x_bits = np.ascontiguousarray(x_bits).view(np.uint32).reshape(-1)
    node_len = np.ascontiguousarray(node_len).view(np.uint32).reshape(-1)
    W = int(n.shape[0])
    th = _host_threads(threads)
    flags = (_native.COMPACT_PAIRS | (_native.COMPACT_REPLICATE if replicate else 0)) if pairs else 0
    m_out = np.zeros(W, dtype=np.int32)
    runs = np.zeros(W, dtype=np.int64)
    rc = L.impop_compact_scan(W, n.ctypes.data, m.ctypes.data, pitch_words.ctypes.data, x_off.ctypes.data,
                              len_off.ctypes.data, x_bits.ctypes.data, node_len.ctypes.data, th, flags, m_out.ctypes.data,
                              runs.ctypes.data)
    compact_batch.last_site_runs = runs              # variant sites (bubble-like runs) of each window in its original node order
    if rc:
        raise NativeError(rc, "impop_compact_scan")
    pitch_out = np.maximum(4, ((m_out + 127) // 128) * 4).astype(np.int32)
    if uniform_pitch and W:
        pitch_out[:] = pitch_out.max()
    rows = n.astype(np.int64) * pitch_out
    x_off_out = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
    cols = pitch_out.astype(np.int64) * 32
    len_off_out = np.concatenate([[0], np.cumsum(cols)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
    x_out = np.zeros(int(rows.sum()) if W else 0, dtype=np.uint32)
    len_out = np.zeros(int(cols.sum()) if W else 0, dtype=np.uint32)
    out = Compacted(m_out, pitch_out, x_off_out, len_off_out, x_out, len_out, runs)
    if pairs:
        n64 = n.astype(np.int64)
        out.row_off_out = np.concatenate([[0], np.cumsum(n64)[:-1]]).astype(np.int64) if W else np.zeros(0, np.int64)
        out.row_adj = np.zeros(int(n64.sum()) if W else 0, dtype=np.int32)
        out.win_const = np.zeros(W, dtype=np.int64)
        out.col_mult = np.zeros(len_out.shape[0], dtype=np.uint8)
    ptr = lambda a: a.ctypes.data if a is not None else None
    rc = L.impop_compact_fill(W, n.ctypes.data, m.ctypes.data, pitch_words.ctypes.data, x_off.ctypes.data,
                              len_off.ctypes.data, x_bits.ctypes.data, node_len.ctypes.data, th, flags, pitch_out.ctypes.data,
                              x_off_out.ctypes.data, len_off_out.ctypes.data, x_out.ctypes.data, len_out.ctypes.data,
                              ptr(out.row_off_out), ptr(out.row_adj), ptr(out.win_const), ptr(out.col_mult))
    if rc:
        raise NativeError(rc, "impop_compact_fill")
    return out


@dataclass
class CompactedUniform:
    """compact_uniform's result: same-shape windows.  Build the batch with
    WindowBatch.from_uniform(ctx, c.x, c.node_len, labels, L, **c.batch_kwargs())."""
    x: np.ndarray            # [W, n, pitch_out] uint32
    node_len: np.ndarray     # [W, 32 * pitch_out] uint32
    m: np.ndarray            # [W] columns in use
    site_runs: np.ndarray    # [W] variant sites counted on the original node order
    row_adj: np.ndarray | None = None     # [W, n] int32
    win_const: np.ndarray | None = None   # [W] int64
    col_mult: np.ndarray | None = None    # [W, 32 * pitch_out] uint8

    def batch_kwargs(self, lo: int = 0, hi: int | None = None, upload=None) -> dict:
        """Keyword arguments of WindowBatch.from_uniform for windows [lo, hi) beyond x / node_len; `upload` (host array ->
        device array) is applied to the arrays the batch reads on the device (default: from_uniform uploads them)."""
        sl = slice(lo, hi)
        up = upload or (lambda a: a)
        kw = {"site_runs": self.site_runs[sl], "m": self.m[sl], "heavy_entries": heavy_entries(self.node_len[sl])}
        if self.row_adj is not None:
            kw.update(row_adj=up(self.row_adj[sl]), win_const=self.win_const[sl], col_mult=up(self.col_mult[sl]))
        return kw


def heavy_entries(node_len: np.ndarray) -> np.ndarray:
    """Heavy-table entries per window of node lengths [W, m]: sum of ceil(floor(len / 255) / 255) (impop_batch_desc_t
    heavy_entries_host: with it batch set-up never reads the lengths on the host)."""
    l = np.asarray(node_len).astype(np.int64)
    return ((l // 255 + 254) // 255).sum(axis=-1).astype(np.int32)


def compact_uniform(x_bits: np.ndarray, node_len: np.ndarray, threads=None, pairs: bool = True, replicate: bool = True) -> CompactedUniform:
    """Same-shape windows x_bits [W, n, pitch] / node_len [W, m_pad] (host uint32) -> compacted same-shape windows
    (x [W, n, pitch_out], node_len [W, 32 * pitch_out], m [W] + the affine arrays); pitch_out fits the widest compacted window."""
    W, n, pitch = x_bits.shape
    m_pad = node_len.shape[1]
    ar = np.arange(W, dtype=np.int64)
    c = compact_batch(np.full(W, n), np.full(W, m_pad), np.full(W, pitch), ar * (n * pitch), ar * m_pad, x_bits, node_len, threads,
                      uniform_pitch=True, pairs=pairs, replicate=replicate)
    po = int(c.pitch_out[0]) if W else 4
    compact_uniform.last_site_runs = c.site_runs
    out = CompactedUniform(c.x_out.reshape(W, n, po), c.len_out.reshape(W, po * 32), c.m_out, c.site_runs)
    if pairs:
        out.row_adj, out.win_const, out.col_mult = c.row_adj.reshape(W, n), c.win_const, c.col_mult.reshape(W, po * 32)
    return out


def compact_window(win: GraphWindow, pairs: bool = True, replicate: bool = True) -> GraphWindow:
    """One GraphWindow -> its compacted form (visit counts are not carried over: expand multisets first)."""
    c = compact_batch([win.n], [win.m], [win.x_bits.shape[1]], [0], [0], win.x_bits, win.node_len, threads=1, pairs=pairs,
                      replicate=replicate)
    mo, po = int(c.m_out[0]), int(c.pitch_out[0])
    out = GraphWindow(list(win.names), c.x_out.reshape(win.n, po), c.len_out[:mo].copy(), None, win.region, win.length)
    out.site_runs = int(c.site_runs[0])     # counted on the original node order (IMPOP_ST_S_BUBBLES)
    if pairs:
        out.row_adj, out.win_const, out.col_mult = c.row_adj.copy(), int(c.win_const[0]), c.col_mult[:mo].copy()
    return out


def compact_windows(windows, threads=None, pairs: bool = True, replicate: bool = True) -> list:
    """A list of GraphWindows -> their compacted forms, all windows in ONE native call that spreads them over the host
    threads (compact_window one by one is a Python loop on one core: seconds per chromosome)."""
    windows = list(windows)
    if not windows:
        return []
    n = np.array([w.n for w in windows], dtype=np.int32)
    m = np.array([w.m for w in windows], dtype=np.int32)
    pitch = np.array([w.x_bits.shape[1] for w in windows], dtype=np.int32)
    xs = n.astype(np.int64) * pitch
    x_off = np.concatenate([[0], np.cumsum(xs)[:-1]]).astype(np.int64)
    len_off = np.concatenate([[0], np.cumsum(m.astype(np.int64))[:-1]]).astype(np.int64)
    x = np.concatenate([np.ascontiguousarray(w.x_bits, dtype=np.uint32).reshape(-1) for w in windows]) if xs.sum() else np.zeros(4, np.uint32)
    nl = np.concatenate([np.asarray(w.node_len, dtype=np.uint32) for w in windows]) if m.sum() else np.zeros(4, np.uint32)
    c = compact_batch(n, m, pitch, x_off, len_off, x, nl, threads=threads, pairs=pairs, replicate=replicate)
    out = []
    for k, w in enumerate(windows):
        mo, po, nn = int(c.m_out[k]), int(c.pitch_out[k]), int(n[k])
        xo, lo = int(c.x_off_out[k]), int(c.len_off_out[k])
        g = GraphWindow(list(w.names), c.x_out[xo:xo + nn * po].reshape(nn, po), c.len_out[lo:lo + mo], None, w.region, w.length)
        g.site_runs = int(c.site_runs[k])
        if pairs:
            ro = int(c.row_off_out[k])
            g.row_adj, g.win_const, g.col_mult = c.row_adj[ro:ro + nn], int(c.win_const[k]), c.col_mult[lo:lo + mo]
        out.append(g)
    return out
