#!/usr/bin/env python3
"""Pipeline model of the pairs kernel: the producer -> MMA chain fills one of NBUF TMEM accumulators per
item, the three epilogue warps of each TMEM lane quarter drain it; an accumulator is free again when ALL four quarters
are done with it.  Unit = one 16-column chunk of one epilogue warp (about 4.5 k cycles measured).  Prints the time per
466-haplotype window for item orders, accumulator counts and chain speeds.  No GPU needed.

    python tools/pipeline_model.py [n_haplotypes]
"""
import itertools
import sys

import numpy as np

n = int(sys.argv[1]) if len(sys.argv) > 1 else 466


def items(n, cap):
    """(row block, first column, columns) as impop_batch_create cuts them (common.cuh: items_of_rowblock)."""
    out = []
    for bi in range((n + 127) // 128):
        width = n - 128 * bi
        parts = (width + cap - 1) // cap
        per = ((width + parts - 1) // parts + 15) // 16 * 16
        for p in range(parts):
            c0 = 128 * bi + p * per
            w = min(per, ((n - c0) + 15) // 16 * 16)
            if w > 0:
                out.append((bi, c0, w))
    return out


def loads(its):
    """Valid 16-column chunks per TMEM lane quarter and item (every other diagonal block is stored reversed)."""
    W = []
    for bi, col0, nc in its:
        rev = col0 == 128 * bi and bi % 2 == 1
        w = [0] * 4
        for rq in range(4):
            r0 = 128 * bi + 32 * rq
            if r0 < n:
                width = min(nc, n - col0)
                chi = (width + 15) // 16
                clo = min(max(0, (r0 - col0) >> 4), chi)
                w[3 - rq if rev else rq] = chi - clo
        W.append(w)
    return np.array(W, float)


def simulate(W, chain, nbuf=2, warps_per_quarter=3, reps=40):
    Wr, M = np.tile(W, (reps, 1)), np.tile(chain, reps)
    qfree, done, t_mma = np.zeros(4), np.zeros((len(Wr), 4)), 0.0
    for k in range(len(Wr)):
        start = t_mma if k < nbuf else max(t_mma, done[k - nbuf].max())
        t_mma = start + M[k]
        for q in range(4):
            s = max(qfree[q], t_mma)
            done[k, q] = qfree[q] = s + Wr[k, q] / warps_per_quarter
    return done.max() / reps


for cap, nbuf in ((256, 2), (256, 3), (160, 3)):
    its = items(n, cap)
    W = loads(its)
    ideal = W.sum() / 12.0
    for per_item in (3.0, 2.6, 2.2):                    # chain time of a 128 x 240 item in epilogue chunk units
        chain = np.array([per_item / 368.0 * (128 + w) for _, _, w in its])
        print(f"cap {cap} accumulators {nbuf} items {len(its)}: chain {per_item:.1f} -> {simulate(W, chain, nbuf):6.2f} units per window "
              f"(ideal {ideal:.2f}, chain alone {chain.sum():.1f})")
its = items(n, 256)
W = loads(its)
if len(its) <= 7:
    chain = np.array([2.6 / 368.0 * (128 + w) for _, _, w in its])
    best = min(((simulate(W[list(p)], chain[list(p)]), p) for p in itertools.permutations(range(len(its)))), key=lambda t: t[0])
    print(f"item order: shipped {simulate(W, chain):.2f}, best permutation {best[0]:.2f} {best[1]}")
