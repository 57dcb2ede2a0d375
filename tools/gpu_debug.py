#!/usr/bin/env python3
"""Bring-up diagnostics for the tcgen05 pairwise kernel: structured inputs whose intersections
reveal a wrong operand layout (which rows / K slabs went where), compared with SIMT and the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from impop_b200.engine import Context, WindowBatch  # noqa: E402
from oracle import clib, similarity  # noqa: E402

ctx = Context(0)
rng = np.random.default_rng(0)


def run(name, x, node_len):
    n, m = x.shape
    bits = similarity.pack_bits(x)
    batch = WindowBatch.from_windows(ctx, [(bits, node_len, np.full(n, 15, dtype=np.uint8), 1000)])
    A0, I0, pi0 = clib.window_pairwise(bits, m, node_len)
    for algo, an in ((1, "simt"), (0, "tc")):
        try:
            I, A, pi = batch.pairwise(0, algo)
            ctx.check()
        except Exception as exc:
            print(f"[{name}] {an}: ERROR {exc}")
            continue
        I, A, pi = I.cpu().numpy(), A.cpu().numpy(), pi.cpu().numpy()
        bad = np.argwhere(I != I0)
        print(f"[{name}] {an}: A ok={bool((A == A0).all())} I mismatches={len(bad)}/{n * n} pi exact={bool((pi == pi0).all())}")
        if len(bad):
            for i, j in bad[:8]:
                print(f"    I[{i},{j}] got {I[i, j]} want {I0[i, j]}")
            print("    rows with mismatch:", np.unique(bad[:, 0])[:20], " cols:", np.unique(bad[:, 1])[:20])
    batch.close()


# 1. one-hot: haplotype i carries node i only, len = i + 1  -> I is diagonal with i + 1
n = 128
x = np.eye(n, dtype=np.uint8)
run("onehot128", x, np.arange(1, n + 1, dtype=np.uint32))
# 2. all ones, unit lengths: I = m everywhere
run("ones", np.ones((64, 96), dtype=np.uint8), np.ones(96, dtype=np.uint32))
# 3. random, small lens, n = 256 (one full 128 x 256 item)
x = (rng.random((256, 200)) < 0.5).astype(np.uint8)
run("rand256", x, rng.integers(0, 200, size=200).astype(np.uint32))
# 4. heavy lengths
run("heavy", x[:100], rng.integers(255, 100000, size=200).astype(np.uint32))
# 5. 466 x 1024
x = (rng.random((466, 1009)) < 0.6).astype(np.uint8)
run("hprc", x, rng.integers(1, 300, size=1009).astype(np.uint32))
print("launches", ctx.launches)
