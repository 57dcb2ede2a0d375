"""All-pairs similarity tables (the text hand-off the reference scripts consume), kept as a
dense identity matrix so the reductions run on the device.

The reference scripts hold the table as ``dict[(min_name, max_name)] -> float``
(pica2.py:6-58, h-fst.py:84-119).  `SimilarityTable` offers the same mapping interface
(so code written against the reference's return value keeps working) and additionally owns
the dense n x n fp64 matrix (NaN = pair absent) that libimpop_b200 reduces on the GPU.
Only parsing and name handling happen here; no statistic is computed on the host.
"""
from __future__ import annotations

import csv
from collections.abc import Mapping

import numpy as np

REQUIRED_COLUMNS = ("group.a", "group.b", "estimated.identity")


class SimilarityTable(Mapping):
    """Mapping (name_a, name_b) -> estimated.identity backed by a dense matrix over sorted names."""

    def __init__(self, names, matrix: np.ndarray):
        self.names = list(names)
        self.index = {s: i for i, s in enumerate(self.names)}
        self.matrix = matrix
        self._device = {}          # (context, round_digits) -> device array

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_rows(cls, rows, combine: str = "last"):
        """rows: iterable of (a, b, value).  combine='last' (pica2.py:44 / h-fst.py:112: a repeated pair
        keeps the last value) or 'max' (af.py:37-39 links on ANY row reaching the threshold)."""
        rows = list(rows)
        names = sorted({r[0] for r in rows} | {r[1] for r in rows})
        index = {s: i for i, s in enumerate(names)}
        n = len(names)
        mat = np.full((n, n), np.nan, dtype=np.float64)
        if rows:
            ia = np.fromiter((index[r[0]] for r in rows), dtype=np.int64, count=len(rows))
            ib = np.fromiter((index[r[1]] for r in rows), dtype=np.int64, count=len(rows))
            v = np.fromiter((r[2] for r in rows), dtype=np.float64, count=len(rows))
            if combine == "max":
                # later (larger) assignments win; a literal `nan` row never hides a real one (af.py:38 links on ANY row that
                # reaches the threshold, and nan >= t is False): NaNs first, then the finite values in ascending order
                order = np.argsort(np.where(np.isnan(v), -np.inf, v), kind="stable")
                ia, ib, v = ia[order], ib[order], v[order]
            lo, hi = np.minimum(ia, ib), np.maximum(ia, ib)
            mat[lo, hi] = v                                # repeated index: the last assignment stays
            mat = np.where(np.isnan(mat), mat.T, mat)      # mirror the upper triangle
        return cls(names, mat)

    @classmethod
    def from_mapping(cls, similarity, elements=None):
        """Accept what the reference's read_similarity_file returns: a dict keyed by name pairs."""
        if isinstance(similarity, cls):
            return similarity
        rows = [(a, b, float(v)) for (a, b), v in similarity.items()]
        tab = cls.from_rows(rows)
        if elements is not None:
            extra = sorted(set(elements) - set(tab.names))
            if extra:                                       # elements that occur in no pair
                names = sorted(tab.names + extra)
                mat = np.full((len(names), len(names)), np.nan)
                where = [names.index(s) for s in tab.names]
                mat[np.ix_(where, where)] = tab.matrix
                tab = cls(names, mat)
        return tab

    # ------------------------------------------------------------------ Mapping interface
    def _key(self, key):
        a, b = key
        i, j = self.index.get(a), self.index.get(b)
        if i is None or j is None:
            return None
        return (i, j)

    def __getitem__(self, key):
        ij = self._key(key)
        if ij is None:
            raise KeyError(key)
        v = self.matrix[ij]
        if v != v:
            raise KeyError(key)
        return float(v)

    def __iter__(self):
        n = len(self.names)
        for i in range(n):
            for j in range(i, n):
                if self.matrix[i, j] == self.matrix[i, j]:
                    yield (self.names[i], self.names[j])

    def __len__(self):
        iu = np.triu_indices(len(self.names))
        return int(np.count_nonzero(~np.isnan(self.matrix[iu])))

    # ------------------------------------------------------------------ device side
    def rounded(self, digits) -> np.ndarray:
        """Host restatement of round(sim, r) (one CPython round() per element): only for `digits` the device kernel does
        not take (negative, or beyond 22 where 10^r is no longer exact)."""
        if digits is None:
            return self.matrix
        flat = [v if v != v else round(v, digits) for v in self.matrix.ravel().tolist()]
        return np.array(flat, dtype=np.float64).reshape(self.matrix.shape)

    def device(self, ctx, round_digits=None):
        """The (optionally rounded) matrix as a device array, uploaded once per context.  round(sim, r) as the reference
        applies it to every value (pica2.py:81-83, h-fst.py:149-150) runs on the device (impop_round_decimal: exact CPython
        semantics -- correctly rounded decimal, which rint(x * 10^r) / 10^r is not)."""
        key = (id(ctx), round_digits)
        if key not in self._device:
            on_device = round_digits is not None and 0 <= round_digits <= 22
            host = self.matrix if (round_digits is None or on_device) else self.rounded(round_digits)
            host = np.ascontiguousarray(host)
            if host.size == 0:
                host = np.zeros((0, 0), dtype=np.float64)
            dev = ctx.upload(host)
            if on_device and host.size:
                ctx.round_decimal(dev, round_digits)
            self._device[key] = dev
        return self._device[key]

    def host(self, ctx, round_digits=None) -> np.ndarray:
        """The (optionally rounded) matrix back on the host (sub-matrix selection of hud.py's grouped method)."""
        if round_digits is None:
            return self.matrix
        key = (id(ctx), round_digits, "host")
        if key not in self._device:
            self._device[key] = self.device(ctx, round_digits).cpu().numpy()
        return self._device[key]

    def labels(self, ctx, **classes):
        """uint8 label vector on the device from name collections: subset=..., a=..., b=..., seg=..."""
        bits = {"subset": 1, "a": 2, "b": 4, "seg": 8}
        lab = np.zeros(len(self.names), dtype=np.uint8)
        for cls_name, members in classes.items():
            if members is None:
                continue
            idx = [self.index[s] for s in members if s in self.index]
            lab[idx] |= bits[cls_name]
        return ctx.upload(lab)


def read_table_fast(path):
    """(SimilarityTable, data rows) through libimpop_b200's native reader (impop_tsv_scan / impop_tsv_fill), or None
    when the text is not machine-clean (quotes, non-ASCII, short rows, unusual numbers, missing columns, empty
    file ...) -- the caller then runs `read_rows`, which mirrors the reference's handling of such input."""
    import ctypes as C

    from . import _native
    try:
        with open(path, "rb") as fh:
            text = fh.read()
    except OSError:
        return None
    lib = _native.lib()
    info = _native.TsvInfo()
    if lib.impop_tsv_scan(text, len(text), C.byref(info)) != 0 or info.status != 0:
        return None
    n = int(info.names)
    mat = np.empty((n, n), dtype=np.float64)
    names_buf = np.zeros(max(int(info.name_bytes), 1), dtype=np.uint8)
    name_off = np.zeros(n + 1, dtype=np.int64)
    if lib.impop_tsv_fill(text, len(text), mat.ctypes.data, names_buf.ctypes.data, name_off.ctypes.data) != 0:
        return None
    raw = names_buf.tobytes()
    names = [raw[name_off[i]:name_off[i + 1] - 1].decode("ascii") for i in range(n)]
    return SimilarityTable(names, mat), int(info.rows)


class TableFormatError(Exception):
    """A similarity table the reference scripts would refuse (message mirrors theirs)."""


def read_rows(handle, on_bad_value: str = "error", strip_coords: bool = False):
    """Parse a tab-separated table with a header; columns are found by name, extras ignored.

    on_bad_value: 'error' (pica2.py:37-41), 'skip' (h-fst.py:107-109) or 'raise' (af.py:15, a bare
    float() -> ValueError).  Returns (rows, row_count, bad) where bad lists (line_number, text)."""
    reader = csv.DictReader(handle, delimiter="\t")
    if not reader.fieldnames:
        raise TableFormatError("empty")
    missing = set(REQUIRED_COLUMNS) - set(reader.fieldnames)
    if missing:
        raise TableFormatError(("columns", list(reader.fieldnames)))
    rows, bad, count = [], [], 0
    for line_no, rec in enumerate(reader, start=2):
        count += 1
        a, b, text = rec["group.a"], rec["group.b"], rec["estimated.identity"]
        try:
            v = float(text)
        except (TypeError, ValueError):
            if on_bad_value == "raise":
                raise
            bad.append((line_no, text))
            if on_bad_value == "error":
                break
            continue
        if strip_coords:
            a, b = a.split(":", 1)[0], b.split(":", 1)[0]
        rows.append((a, b, v))
    return rows, count, bad
