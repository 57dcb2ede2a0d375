"""Tolerance policy for comparing device statistics rows with oracle rows.  TEST INFRASTRUCTURE ONLY.

North star: intersections / unions / S / counts bit-exact; pi, Fst, D within 1e-12 relative in fp64.
"Relative" needs a scale for the three columns that are DIFFERENCES of nearly equal fp64 numbers:

  da  = dxy - pi_xy          scale = max(|dxy|, |pi_xy|)
  fst = da / dxy             scale = max(1, |pi_xy / dxy|)          (= the da scale divided by dxy)
  D   = (pi - S/a1) / den    scale = max(|pi|, S/a1) / den

because the reference's own value of such a difference moves by (ulp of the operands) x (condition
number) when the same table is summed in another order (its sets iterate in hash order).  Every
other column is compared at 1e-12 relative to the value itself.  Both sides sum with compensation
(CPython's sum() / two-sum on the device), so in practice the columns agree to a few ulp.
"""
from __future__ import annotations

import math

TOL = 1e-12
COLS = ("pi", "pi_per_site", "pi_a", "pi_b", "pi_xy", "dxy", "da", "fst", "S", "tajima_d", "a1", "e1", "e2", "n",
        "sum_S", "sum_AA", "sum_BB", "sum_AB", "tajima_d_raw", "S_bubbles")


def _close(a, b, scale, tol):
    if a != a or b != b:
        return a != a and b != b
    if a == b:
        return True
    if math.isinf(a) or math.isinf(b):
        return False
    return abs(a - b) <= tol * scale


def row_mismatches(got, want, tol: float = TOL, sites: bool = False):
    """List of (column name, got, want, allowed) for the columns of one 20-wide statistics row that disagree.
    `sites`: also compare S_bubbles (site runs depend on the node ORDER: only comparable when both sides saw the same
    column order, i.e. not between an ingest-compacted batch and its original)."""
    got = [float(v) for v in got]
    want = [float(v) for v in want]
    bad = []
    dxy, pi_xy = abs(want[5]), abs(want[4])
    S, a1, e1, e2 = want[8], want[10], want[11], want[12]
    den = math.sqrt(e1 * S + e2 * S * (S - 1.0)) if (S > 0 and a1 == a1 and e1 * S + e2 * S * (S - 1.0) > 0) else float("nan")
    for k, name in enumerate(COLS):
        g, w = got[k], want[k]
        if name == "S_bubbles" and not sites:
            continue
        if name == "da":
            scale = max(dxy, pi_xy, abs(w))
        elif name == "fst":
            scale = max(1.0, (pi_xy / dxy) if dxy > 0 else 1.0, abs(w))
        elif name in ("tajima_d", "tajima_d_raw") and den == den and den > 0:
            pi_used = want[0] if (name == "tajima_d_raw" or want[1] != want[1]) else want[1]
            scale = max(abs(w), max(abs(pi_used), S / a1) / den)
        else:
            scale = max(abs(g), abs(w))
        if not _close(g, w, scale, tol):
            bad.append((name, g, w, tol * scale))
    return bad


def rows_close(got, want, tol: float = TOL, sites: bool = False):
    """(ok, first mismatch description) over 2-D arrays of rows."""
    for r in range(len(want)):
        bad = row_mismatches(got[r], want[r], tol, sites)
        if bad:
            return False, f"row {r}: {bad[:3]}"
    return True, ""
