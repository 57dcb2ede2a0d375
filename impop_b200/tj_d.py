"""Tajima's D from (n, S, pi) -- device-backed drop-in for the reference's scripts/tj_d.py
(tj_d.py:28-88: same dataclass, function signature, ValueErrors, CLI flags and stdout).

The harmonic sums a1, a2 (formed as CPython >= 3.12's compensated `sum()` forms them) and the
closed form run in libimpop_b200 (impop_tajima_d); `tajimas_d_batch` evaluates many windows in
one launch.  No CPU fallback.
"""
from __future__ import annotations

import argparse
import sys
from dataclasses import dataclass

import numpy as np

from .runtime import default_context


@dataclass
class TajimaComponents:
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


def tajimas_d_batch(n, S, pi, ctx=None, with_parts=False):
    """Vectorised tj_d.tajimas_d: sequences of n (int), S, pi (float) -> list of D (and component rows)."""
    ctx = ctx or default_context()
    n, S, pi = list(n), list(S), list(pi)
    for nn, ss, pp in zip(n, S, pi):
        if nn < 2:
            raise ValueError("n must be >= 2")
        if ss < 0 or pp < 0:
            raise ValueError("S and pi must be non-negative")
    nt = ctx.upload(np.asarray(n, dtype=np.int64))
    st = ctx.upload(np.asarray(S, dtype=np.float64))
    pt = ctx.upload(np.asarray(pi, dtype=np.float64))
    out = ctx.tajima_d(nt, st, pt, with_parts=with_parts)
    ctx.check()
    if with_parts:
        return out[0].cpu().tolist(), out[1].cpu().tolist()
    return out.cpu().tolist()


def tajimas_d(n: int, S: float, pi: float, return_components: bool = False, ctx=None):
    """D = (pi - S/a1) / sqrt(e1*S + e2*S*(S-1)); NaN when S == 0 or the denominator is 0 (tj_d.py:47-69)."""
    if n < 2:
        raise ValueError("n must be >= 2")
    if S < 0 or pi < 0:
        raise ValueError("S and pi must be non-negative")
    d, parts = tajimas_d_batch([int(n)], [float(S)], [float(pi)], ctx=ctx, with_parts=True)
    if return_components:
        return d[0], TajimaComponents(*parts[0])
    return d[0]


def main(argv=None):
    parser = argparse.ArgumentParser(description="Compute Tajima's D from n, S, and pi.")
    parser.add_argument("-n", "--sample-size", type=int, required=True, help="Number of sequences (n >= 2)")
    parser.add_argument("-S", "--segregating-sites", type=float, required=True, help="Number of segregating sites S (>= 0)")
    parser.add_argument("-p", "--pi", type=float, required=True, help="Mean pairwise differences pi (>= 0)")
    parser.add_argument("--show-components", action="store_true", help="Print intermediate constants (a1, a2, e1, e2, etc.)")
    args = parser.parse_args(argv)
    D, comps = tajimas_d(args.sample_size, args.segregating_sites, args.pi, return_components=True)
    print(f"Tajima's D: {D}")
    if args.show_components:
        print("--- Components ---")
        print(f"a1={comps.a1} a2={comps.a2}")
        print(f"b1={comps.b1} b2={comps.b2}")
        print(f"c1={comps.c1} c2={comps.c2}")
        print(f"e1={comps.e1} e2={comps.e2}")
        print(f"numerator={comps.numerator} denominator={comps.denominator}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
