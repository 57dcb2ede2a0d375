"""Rows a-1 .. a-10: array-based CPU restatement of the reference's statistics scripts.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against the unmodified
reference scripts through tests/golden/*.json (made by tests/golden/make_golden.py).

The reference keeps the pair table in a dict keyed by name tuples and sums in
set-iteration order; this restatement keeps a dense n x n fp64 matrix (NaN =
pair absent from the table) indexed by the sorted name list.  Every per-pair
term is formed with the same fp64 operations as the reference, so results differ
from the reference only by summation order (observed <= 1e-15 relative).

Citations are to /root/reference/scripts.
"""
from __future__ import annotations

import csv
import math
from dataclasses import dataclass

import numpy as np


# ----------------------------------------------------------------------------
# a-1 / a-4 : similarity table -> dense matrix
# ----------------------------------------------------------------------------
def neumaier_sum(values) -> float:
    """Python >= 3.12 `sum()` over floats: Neumaier compensated summation (CPython
    Python/bltinmodule.c, builtin_sum_impl).  The reference calls the builtin
    (tj_d.py:41-45, pica2.py:154, h-fst.py:171) and the golden vectors were made on
    CPython 3.12.3, so the restatement spells the algorithm out instead of relying
    on the interpreter it happens to run under."""
    total = 0.0
    comp = 0.0
    for x in values:
        t = total + x
        if abs(total) >= abs(x):
            comp += (total - t) + x
        else:
            comp += (x - t) + total
        total = t
    if comp and math.isfinite(comp):
        total += comp
    return total


class TableError(Exception):
    """Raised where the reference prints a message and exits with status 1."""


def parse_similarity_tsv(path, strict: bool = True, strip_coords: bool = False):
    """TSV -> (names sorted, dense identity matrix with NaN for absent pairs, row count).

    pica2.py:6-58 (strict=True: a bad float aborts, :37-41) and h-fst.py:84-119
    (strict=False: a bad float is skipped with a warning, :107-109).  Columns are
    located by header name (csv.DictReader); a repeated pair keeps the last value
    (pica2.py:44).  strip_coords=True applies af.py:13-14 (name cut at first ':').
    """
    try:
        fh = open(path, newline="")
    except FileNotFoundError as exc:
        raise TableError(f"File not found {path}") from exc
    with fh:
        reader = csv.DictReader(fh, delimiter="\t")
        if not reader.fieldnames:
            raise TableError(f"File {path} is empty or missing a header")
        need = {"group.a", "group.b", "estimated.identity"}
        if not need.issubset(reader.fieldnames):
            raise TableError(f"File must contain columns: {sorted(need)}")
        rows = []
        nrows = 0
        for line_no, rec in enumerate(reader, start=2):
            nrows += 1
            a, b = rec["group.a"], rec["group.b"]
            try:
                v = float(rec["estimated.identity"])
            except (TypeError, ValueError):
                if strict:
                    raise TableError(f"Invalid similarity value on line {line_no}")
                continue
            if strip_coords:
                a, b = a.split(":", 1)[0], b.split(":", 1)[0]
            rows.append((a, b, v))
    names = sorted({r[0] for r in rows} | {r[1] for r in rows})
    index = {s: i for i, s in enumerate(names)}
    mat = np.full((len(names), len(names)), np.nan)
    for a, b, v in rows:
        i, j = index[a], index[b]
        mat[i, j] = v
        mat[j, i] = v
    return names, mat, nrows


def py_round_matrix(mat: np.ndarray, digits):
    """round(sim, r) exactly as CPython does it (pica2.py:81-83, h-fst.py:149-150)."""
    if digits is None:
        return mat
    flat = [v if v != v else round(v, digits) for v in mat.ravel().tolist()]
    return np.array(flat, dtype=np.float64).reshape(mat.shape)


# ----------------------------------------------------------------------------
# a-2 : pica2 nucleotide diversity
# ----------------------------------------------------------------------------
def greedy_groups(mat: np.ndarray, names, threshold: float):
    """Star grouping of pica2.py:94-112 made deterministic.

    The reference pops an arbitrary element of a set of strings as the seed
    (pica2.py:100), which depends on PYTHONHASHSEED.
    Here the seed is always the smallest remaining name.  The two agree whenever
    "sim > threshold" is an equivalence relation on the table, and trivially when
    threshold >= every identity (each haplotype its own group).  Returns a sorted
    list of sorted index lists (groups.sort(), pica2.py:112).
    """
    n = len(names)
    free = np.ones(n, dtype=bool)
    groups = []
    for seed in range(n):               # names are sorted, so index order == name order
        if not free[seed]:
            continue
        free[seed] = False
        row = mat[seed]
        take = free & (row > threshold)  # NaN > t is False: absent pair never joins (:106)
        members = [seed] + np.nonzero(take)[0].tolist()
        free[take] = False
        groups.append(sorted(members))
    groups.sort()
    return groups


def pica2_pi(mat: np.ndarray, names, threshold: float = 1.0, sequence_length=None, round_digits=None):
    """(pi, pi_per_site) as pica2.analyze_similarity_matrix returns them (pica2.py:60-169)."""
    mat = py_round_matrix(mat, round_digits)
    groups = greedy_groups(mat, names, threshold)
    total = sum(len(g) for g in groups)
    if total == 0:
        return 0.0, 0.0                                   # pica2.py:122-124
    terms = []
    for gi in range(len(groups)):
        for gj in range(gi + 1, len(groups)):
            s = float(mat[groups[gi][0], groups[gj][0]])   # representatives, :128
            if s != s:                                     # absent -> skipped, :132-134
                continue
            fi = len(groups[gi]) / total
            fj = len(groups[gj]) / total
            terms.append((1 - s) * fi * fj)                # :137-139
    if not terms:
        return 0.0, 0.0                                   # :150-152
    pi = (total / (total - 1)) * neumaier_sum(2 * t for t in terms)  # :154 (n == 1 -> ZeroDivisionError)
    per_site = pi / sequence_length if sequence_length else None  # :161-164
    return pi, per_site


# ----------------------------------------------------------------------------
# a-3 : population lists -> sequence names
# ----------------------------------------------------------------------------
_HAP_SUFFIX = (("_hap1", "#1#"), ("_hap2", "#2#"), ("_mat", "#1#"), ("_pat", "#2#"))


def canonical_prefix(identifier: str) -> str:
    """h-fst.py:18-61: assembly name -> PanSN prefix used with str.startswith."""
    if not identifier:
        return ""
    tok = identifier.strip()
    if not tok or tok[0] == "#":
        return ""
    cut = tok.find("_hprc")
    if cut >= 0:
        tok = tok[:cut]
    for suffix, tag in _HAP_SUFFIX:
        if tok.endswith(suffix):
            return tok[: len(tok) - len(suffix)] + tag
    if "#" in tok:
        return tok if tok.endswith("#") else tok + "#"
    return tok + "#"


def expand_population(raw_ids, names):
    """h-fst.py:64-82 -> (set of matched names, list of unmatched raw ids)."""
    hit, miss = set(), []
    for raw in raw_ids:
        pre = canonical_prefix(raw)
        if not pre:
            continue
        found = [s for s in names if s.startswith(pre)]
        if found:
            hit.update(found)
        else:
            miss.append(raw)
    return hit, miss


def read_id_list(path):
    """h-fst.py:121-128: stripped non-blank lines not starting with '#'."""
    with open(path) as fh:
        return {ln.strip() for ln in fh if ln.strip() and not ln.startswith("#")}


# ----------------------------------------------------------------------------
# a-5 / a-6 : Hudson Fst
# ----------------------------------------------------------------------------
def mean_diversity(mat: np.ndarray, idx1, idx2=None, round_digits=None):
    """h-fst.py:130-171 -> (mean of 1 - s over present pairs, count, missing)."""
    idx1 = sorted(idx1)
    vals, missing = [], 0
    if idx2 is None:
        pairs = ((idx1[p], idx1[q]) for p in range(len(idx1)) for q in range(p + 1, len(idx1)))
    else:
        idx2 = sorted(idx2)
        pairs = ((i, j) for i in idx1 for j in idx2)
    for i, j in pairs:
        s = float(mat[i, j])          # python float: np.float64.__round__ is not CPython's round
        if s != s:
            missing += 1
            continue
        if round_digits is not None:
            s = round(s, round_digits)
        vals.append(1 - s)
    if not vals:
        return 0.0, 0, missing
    return neumaier_sum(vals) / len(vals), len(vals), missing


def hudson_fst(mat, names, pop_a, pop_b, sequence_length=None, round_digits=None):
    """h-fst.py:173-249.  pop_a / pop_b are sets of sequence names."""
    both = set(pop_a) & set(pop_b)
    pa = sorted(set(pop_a) - both)
    pb = sorted(set(pop_b) - both)
    where = {s: i for i, s in enumerate(names)}
    ia = [where[s] for s in pa if s in where]
    ib = [where[s] for s in pb if s in where]
    # names absent from the table contribute only "missing" pairs in the reference
    pi_a, cnt_a, _ = mean_diversity(mat, ia, round_digits=round_digits)
    pi_b, cnt_b, _ = mean_diversity(mat, ib, round_digits=round_digits)
    pi_xy = 0.5 * (pi_a + pi_b)
    dxy, cnt_ab, _ = mean_diversity(mat, ia, ib, round_digits=round_digits)
    fst = (dxy - pi_xy) / dxy if dxy > 0 else 0.0
    if sequence_length and sequence_length > 0:
        L = sequence_length
        out = dict(fst=fst, pi_a=pi_a / L, pi_b=pi_b / L, pi_xy=pi_xy / L, dxy=dxy / L, da=(dxy - pi_xy) / L)
    else:
        out = dict(fst=fst, pi_a=pi_a, pi_b=pi_b, pi_xy=pi_xy, dxy=dxy, da=dxy - pi_xy)
    out["counts"] = (cnt_a, cnt_b, cnt_ab)
    return out


def pooled_fst_text(pi_a_txt: str, pi_b_txt: str, pi_c_txt: str):
    """a-10: inline python of run_fst_impg.sh:199-218 on the 8-decimal text pis."""
    pa, pb, pc = float(pi_a_txt), float(pi_b_txt), float(pi_c_txt)
    avg = 0.5 * (pa + pb)
    fst = "NA" if pc == 0 else f"{(pc - avg) / pc:.8f}"
    return f"{avg:.8f}", fst


# ----------------------------------------------------------------------------
# a-7 : Tajima's D
# ----------------------------------------------------------------------------
@dataclass
class TajimaParts:
    a1: float
    a2: float
    b1: float
    b2: float
    c1: float
    c2: float
    e1: float
    e2: float
    numerator: float
    denominator: float


def tajimas_d(n: int, s_sites: float, pi: float):
    """tj_d.py:41-69 -> (D, TajimaParts).  Harmonic sums as builtin sum() forms them."""
    if n < 2:
        raise ValueError("n must be >= 2")
    if s_sites < 0 or pi < 0:
        raise ValueError("S and pi must be non-negative")
    a1 = neumaier_sum(1.0 / i for i in range(1, n))
    a2 = neumaier_sum(1.0 / (i * i) for i in range(1, n))
    b1 = (n + 1.0) / (3.0 * (n - 1.0))
    b2 = 2.0 * (n * n + n + 3.0) / (9.0 * n * (n - 1.0))
    c1 = b1 - (1.0 / a1)
    c2 = b2 - ((n + 2.0) / (a1 * n)) + (a2 / (a1 * a1))
    e1 = c1 / a1
    e2 = c2 / (a1 * a1 + a2)
    num = pi - (s_sites / a1)
    den = math.sqrt(e1 * s_sites + e2 * s_sites * (s_sites - 1.0)) if s_sites > 0 else float("nan")
    d = num / den if (den and den == den and den != 0.0) else float("nan")
    return d, TajimaParts(a1, a2, b1, b2, c1, c2, e1, e2, num, den)


# ----------------------------------------------------------------------------
# a-9 : af.py haplotype clusters
# ----------------------------------------------------------------------------
def af_clusters(mat: np.ndarray, names, threshold: float):
    """af.py:21-54: components of the graph {identity >= threshold}, ordered by
    (-size, sorted members).  Returns list of sorted name lists."""
    n = len(names)
    parent = list(range(n))

    def root(v):
        while parent[v] != v:
            parent[v] = parent[parent[v]]
            v = parent[v]
        return v

    ii, jj = np.nonzero(np.triu(mat >= threshold, k=1))
    for i, j in zip(ii.tolist(), jj.tolist()):
        ri, rj = root(i), root(j)
        if ri != rj:
            parent[max(ri, rj)] = min(ri, rj)
    comps: dict[int, list[str]] = {}
    for v in range(n):
        comps.setdefault(root(v), []).append(names[v])
    return sorted((sorted(c) for c in comps.values()), key=lambda c: (-len(c), c))


def af_summary(clusters):
    """af.py:46-54 -> [(cluster_id, count, frequency, members)]."""
    total = sum(len(c) for c in clusters)
    return [(f"c{k}", len(c), (len(c) / total) if total else 0.0, sorted(c))
            for k, c in enumerate(clusters, 1)]


# ----------------------------------------------------------------------------
# BASELINE config 4 (north star): per-site allele counts / frequencies
# ----------------------------------------------------------------------------
def site_allele_counts(site_bits: np.ndarray, pop_masks: np.ndarray, pop_sizes=None):
    """counts[s, p] = popcount(site_bits[s] & pop_masks[p]);  freq = count / |pop|.

    The only per-site allele-count semantics in the reference is the unfinished
    scripts/wip/op-afs.py:26-45 (count of an allele in a column / column length);
    SURVEY.md note N1.  site_bits: (M, words) uint64/uint32, pop_masks: (P, words).
    """
    sb = np.ascontiguousarray(site_bits).view(np.uint8).reshape(site_bits.shape[0], -1)
    pm = np.ascontiguousarray(pop_masks).view(np.uint8).reshape(pop_masks.shape[0], -1)
    lut = np.array([bin(v).count("1") for v in range(256)], dtype=np.int32)
    counts = np.empty((sb.shape[0], pm.shape[0]), dtype=np.int32)
    for p in range(pm.shape[0]):
        counts[:, p] = lut[sb & pm[p][None, :]].sum(axis=1)
    if pop_sizes is None:
        pop_sizes = lut[pm].sum(axis=1)
    sizes = np.asarray(pop_sizes, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        freq = counts.astype(np.float64) / sizes[None, :]
    freq = np.where(sizes[None, :] > 0, freq, 0.0)
    return counts, freq


# ----------------------------------------------------------------------------
# f-3 : hudson/hud.py, grouped method (deterministic seeds)
# ----------------------------------------------------------------------------
def hud_grouped_diversity(mat: np.ndarray, idx, threshold: float = 0.999):
    """(diversity, groups, missing) of hud.calculate_diversity_grouped (hud.py:99-128) over the rows `idx` (sorted by
    name), with group_sequences (hud.py:64-84) seeded by the smallest remaining name instead of set.pop().
    `mat` is already rounded.  The representative similarity is that of the groups' first members, which is what
    hud.get_group_similarity (hud.py:86-97) finds first on a complete table."""
    idx = list(idx)
    sub = mat[np.ix_(idx, idx)] if idx else np.zeros((0, 0))
    groups = greedy_groups(sub, idx, threshold)
    n_total = len(idx)
    if n_total <= 1:
        return 0.0, len(groups), 0, groups
    total, missing = 0.0, 0                                     # sequential float accumulation as hud.py:108-121
    for i in range(len(groups)):
        for j in range(i + 1, len(groups)):
            s = float(sub[groups[i][0], groups[j][0]])
            if s != s:
                missing += 1
                continue
            total += 2 * (len(groups[i]) / n_total) * (len(groups[j]) / n_total) * (1 - s)
    return total * n_total / (n_total - 1), len(groups), missing, groups


def hud_fst_grouped(mat, names, pop_a, pop_b, sequence_length=None, round_digits=None, threshold: float = 0.999):
    """hud.calculate_fst(method='grouped') (hud.py:172-300)."""
    mat = py_round_matrix(mat, round_digits)
    both = set(pop_a) & set(pop_b)
    where = {s: i for i, s in enumerate(names)}
    ia = [where[s] for s in sorted(set(pop_a) - both) if s in where]
    ib = [where[s] for s in sorted(set(pop_b) - both) if s in where]
    pi_a, _, _, ga = hud_grouped_diversity(mat, ia, threshold)
    pi_b, _, _, gb = hud_grouped_diversity(mat, ib, threshold)
    pi_xy = 0.5 * (pi_a + pi_b)
    dxy = 0.0
    for g1 in ga:
        for g2 in gb:
            s = float(mat[ia[g1[0]], ib[g2[0]]])
            if s != s:
                continue
            dxy += (len(g1) * len(g2)) / (len(ia) * len(ib)) * (1 - s)
    fst = (dxy - pi_xy) / dxy if dxy > 0 else 0.0
    if sequence_length and sequence_length > 0:
        L = sequence_length
        return dict(fst=fst, pi_a=pi_a / L, pi_b=pi_b / L, pi_xy=pi_xy / L, dxy=dxy / L, da=(dxy - pi_xy) / L)
    return dict(fst=fst, pi_a=pi_a, pi_b=pi_b, pi_xy=pi_xy, dxy=dxy, da=dxy - pi_xy)
