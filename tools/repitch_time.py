#!/usr/bin/env python3
"""Device time of impop_repitch_rows on a chr2-sized set of windows (tight 9-11 words -> 12 words per row)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from impop_b200.engine import Context
ctx = Context(0)
W, n = 4854, 466
rng = np.random.default_rng(1)
sp = rng.integers(9, 12, W).astype(np.int32); dp = np.full(W, 12, np.int32); rows = np.full(W, n, np.int32)
so = np.concatenate([[0], np.cumsum(rows.astype(np.int64) * sp)]); do = np.concatenate([[0], np.cumsum(rows.astype(np.int64) * dp)])
src = torch.randint(-2**31, 2**31 - 1, (int(so[-1]),), dtype=torch.int32, device=ctx.torch_device)
dst = torch.empty(int(do[-1]), dtype=torch.int32, device=ctx.torch_device)
for _ in range(3):
    ctx.repitch_rows(src, dst, rows, sp, dp, so[:-1], do[:-1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.repitch_rows(src, dst, rows, sp, dp, so[:-1], do[:-1])
e1.record(); torch.cuda.synchronize()
print("repitch of %d windows: %.3f ms per call, %.1f MB in + %.1f MB out" % (W, e0.elapsed_time(e1) / 10, so[-1] * 4 / 1e6, do[-1] * 4 / 1e6))
