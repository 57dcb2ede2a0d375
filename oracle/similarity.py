"""Row a-0: all-pairs haplotype similarity of one window graph (numpy restatement).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: the
arithmetic belongs to `odgi similarity` / `impg similarity`, which are not in
/root/reference and have no pinned version there.  Call sites in the reference:
scripts/run_pica2_odgi.sh:96, run_pica2_impg.sh:162-168, run_fst_impg.sh:62-67,
run_h-fst.sh:65-67, run_tajd.sh:160, hudson/run_hud.sh:71-73.  The columns the
reference consumes are `group.a`, `group.b`, `estimated.identity`
(scripts/pica2.py:22, h-fst.py:94, af.py:13-15).

Definition restated, for haplotype presence
bits x[i,k] over graph nodes k with node lengths len[k]:

    I_ij = sum_k len_k * x_ik * x_jk          node-length-weighted intersection
    A_i  = sum_k len_k * x_ik                 path length
    U_ij = A_i + A_j - I_ij                   union
    J    = (double)I / (double)U              Jaccard            (1 rounding)
    id   = 2.0*J / (1.0 + J)                  estimated.identity (2J exact, 1+J, divide)
    pi   = 1.0 - id                           what pica2.py:137 / h-fst.py:151 form

The fp64 operation ORDER above is part of the contract:
the device epilogue performs the same correctly-rounded operations.
When U == 0 (two empty paths) the result is defined here as J = 0, id = 0.
"""
from __future__ import annotations

import numpy as np


def path_lengths(x: np.ndarray, node_len: np.ndarray) -> np.ndarray:
    """A_i = sum_k len_k x_ik, exact int64.  x: (n, m) bool/0-1, node_len: (m,)"""
    return x.astype(np.int64) @ node_len.astype(np.int64)


def intersections(x: np.ndarray, node_len: np.ndarray) -> np.ndarray:
    """I_ij for all i, j (n x n, int64, exact).  Diagonal equals A_i."""
    xi = x.astype(np.int64)
    return (xi * node_len.astype(np.int64)[None, :]) @ xi.T


def identity_from_counts(inter: np.ndarray, a: np.ndarray):
    """(U, J, identity, pi) from exact integer I (n x n) and A (n) in the contract op order."""
    inter = inter.astype(np.int64)
    a = a.astype(np.int64)
    union = a[:, None] + a[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        jac = inter.astype(np.float64) / union.astype(np.float64)
    jac = np.where(union == 0, 0.0, jac)
    ident = (2.0 * jac) / (1.0 + jac)
    pi = 1.0 - ident
    return union, jac, ident, pi


def pairwise(x: np.ndarray, node_len: np.ndarray):
    """Full restatement for one window: dict with I, A, U, J, identity, pi (n x n / n)."""
    a = path_lengths(x, node_len)
    inter = intersections(x, node_len)
    union, jac, ident, pi = identity_from_counts(inter, a)
    return {"I": inter, "A": a, "U": union, "J": jac, "identity": ident, "pi": pi}


def segregating_nodes(x: np.ndarray, node_len: np.ndarray, rows=None) -> int:
    """Row a-8 replacement: S = #{k : 0 < sum_{i in rows} x_ik < |rows|, len_k > 0}.

    The reference's S is the number of `povu gfa2vcf` records (run_tajd.sh:148),
    which cannot be reproduced here (povu absent, unpinned); BASELINE.json config 3
    redefines S as segregating nodes.  Parity unpinned at this boundary.
    """
    xs = x if rows is None else x[np.asarray(rows)]
    if xs.shape[0] == 0:
        return 0
    cnt = xs.astype(np.int64).sum(axis=0)
    return int(np.count_nonzero((cnt > 0) & (cnt < xs.shape[0]) & (node_len > 0)))


def site_runs(x: np.ndarray, node_len: np.ndarray, rows=None) -> int:
    """Variant sites as a bubble caller would count them -- the reference's S is the number of `povu gfa2vcf` records of
    the window graph (run_tajd.sh:126-148), one per bubble rather than one per node.  povu is absent and unpinned, so this
    is a stated approximation (parity unpinned): with the nodes in graph order, a site is a maximal run of segregating
    nodes not interrupted by a node EVERY row carries; nodes no row carries and zero-length nodes are transparent."""
    xs = x if rows is None else x[np.asarray(rows)]
    if xs.shape[0] == 0:
        return 0
    cnt = xs.astype(np.int64).sum(axis=0)
    runs, in_run = 0, False
    for k in range(xs.shape[1]):
        if node_len[k] == 0:
            continue
        if cnt[k] == xs.shape[0]:
            in_run = False
        elif cnt[k] > 0 and not in_run:
            runs, in_run = runs + 1, True
    return runs


def pack_bits(x: np.ndarray, pitch_words: int | None = None) -> np.ndarray:
    """Bit-pack rows of a 0/1 matrix into little-endian u32 words.

    Column k of row i is bit (k & 31) of word k >> 5 -- the layout
    include/impop_b200.h documents for the device library.  Rows are padded with
    zero bits to `pitch_words` (default: ceil(m/128)*4, i.e. 16-byte multiples).
    """
    x = np.ascontiguousarray(x.astype(np.uint8))
    n, m = x.shape
    if pitch_words is None:
        pitch_words = ((m + 127) // 128) * 4
    padded = np.zeros((n, pitch_words * 32), dtype=np.uint8)
    padded[:, :m] = x
    packed = np.packbits(padded, axis=1, bitorder="little")
    return packed.view("<u4").reshape(n, pitch_words).copy()


def unpack_bits(words: np.ndarray, m: int) -> np.ndarray:
    n = words.shape[0]
    b = np.unpackbits(np.ascontiguousarray(words).view(np.uint8).reshape(n, -1), axis=1, bitorder="little")
    return b[:, :m]


def write_similarity_tsv(path, names, res, extra_columns: bool = True) -> None:
    """Emit an odgi/impg-style all-pairs table the reference scripts can read.

    Identity is written with repr() (17 significant digits) so the reference sees
    exactly the fp64 values (SURVEY.md section 8 c).  Columns other than group.a /
    group.b / estimated.identity are ignored by the reference (csv.DictReader by name).
    """
    n = len(names)
    ident, inter, a, jac = res["identity"], res["I"], res["A"], res["J"]
    with open(path, "w") as fh:
        if extra_columns:
            fh.write("group.a\tgroup.b\tgroup.a.length\tgroup.b.length\tintersection\t"
                     "jaccard.similarity\testimated.identity\n")
        else:
            fh.write("group.a\tgroup.b\testimated.identity\n")
        for i in range(n):
            for j in range(i + 1, n):
                if extra_columns:
                    fh.write(f"{names[i]}\t{names[j]}\t{int(a[i])}\t{int(a[j])}\t{int(inter[i, j])}\t"
                             f"{float(jac[i, j])!r}\t{float(ident[i, j])!r}\n")
                else:
                    fh.write(f"{names[i]}\t{names[j]}\t{float(ident[i, j])!r}\n")


def compact_columns(x: np.ndarray, node_len: np.ndarray):
    """Restatement of the ingest-time column compaction (impop_compact_scan / _fill, include/impop_b200.h): nodes every
    row visits merge into one node of their summed length, nodes no row visits and nodes of length 0 vanish, the rest is
    ordered by length (stable).  Returns (x', node_len') with the merged node last.  I, A, U and S are unchanged."""
    x = np.asarray(x).astype(np.uint8)
    node_len = np.asarray(node_len).astype(np.int64)
    n = x.shape[0]
    cnt = x.astype(np.int64).sum(axis=0)
    live = node_len > 0
    const = live & (cnt == n) & (n > 0)
    var = live & (cnt > 0) & ~const
    idx = np.flatnonzero(var)
    idx = idx[np.argsort(node_len[idx], kind="stable")]
    c = int(node_len[const].sum())
    cols = [x[:, idx]]
    lens = [node_len[idx]]
    if c > 0:
        cols.append(np.ones((n, 1), dtype=np.uint8))
        lens.append(np.array([c], dtype=np.int64))
    return np.concatenate(cols, axis=1), np.concatenate(lens)


def pairwise_affine(x: np.ndarray, node_len: np.ndarray, row_adj, win_const: int):
    """A window in the affine form of include/impop_b200.h (what impop_compact_fill writes with IMPOP_COMPACT_PAIRS):
    I_ij = sum_k len_k x_ik x_jk + C - R_i - R_j, A_i = I_ii, then the contract's U, J, identity, pi.  The ingest step's
    claim -- checked in tests/test_ingest_cpu.py -- is that this equals pairwise() of the original window bit for bit."""
    r = np.asarray(row_adj).astype(np.int64)
    inter = intersections(x, node_len) + int(win_const) - r[:, None] - r[None, :]
    a = np.diagonal(inter).copy()
    union, jac, ident, pi = identity_from_counts(inter, a)
    return {"I": inter, "A": a, "U": union, "J": jac, "identity": ident, "pi": pi}


def segregating_nodes_affine(x: np.ndarray, node_len: np.ndarray, col_mult, rows=None) -> int:
    """S of a window whose columns stand for col_mult nodes each (merged bubbles: 2, copies of a split weight: 0)."""
    xs = x if rows is None else x[np.asarray(rows)]
    if xs.shape[0] == 0:
        return 0
    cnt = xs.astype(np.int64).sum(axis=0)
    seg = (cnt > 0) & (cnt < xs.shape[0]) & (np.asarray(node_len) > 0)
    return int(np.asarray(col_mult).astype(np.int64)[seg].sum())
