#!/usr/bin/env python3
"""Where the end-to-end time of one step goes: pinned H2D, batch set-up, kernels, D2H, tear-down."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from impop_b200 import synth  # noqa: E402
from impop_b200.engine import Context, WindowBatch, NSTATS, NCOUNTS  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4854
ctx = Context(0)
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 50000, W, seed=0xB201)
lab = np.full(466, 9, dtype=np.uint8); lab[pops == 0] |= 2; lab[pops == 2] |= 4
labels = torch.from_numpy(lab).to(ctx.torch_device)
hx = torch.empty(x.shape, dtype=torch.int32, pin_memory=True); hx.copy_(x)
hl = torch.empty(nl.shape, dtype=torch.int32, pin_memory=True); hl.copy_(nl)
hs = torch.empty((W, NSTATS), dtype=torch.float64, pin_memory=True)
hc = torch.empty((W, NCOUNTS), dtype=torch.int64, pin_memory=True)
dx, dl = torch.empty_like(x), torch.empty_like(nl)
torch.cuda.synchronize()


def tick(label, fn, reps=5):
    ts = []
    out = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{label:32s} min {min(ts) * 1e3:8.3f} ms   median {sorted(ts)[len(ts) // 2] * 1e3:8.3f} ms")
    return out


tick("H2D x (pinned, %d MB)" % (hx.numel() * 4 // 1000000), lambda: dx.copy_(hx, non_blocking=True))
tick("H2D node_len", lambda: dl.copy_(hl, non_blocking=True))
pageable = hx.clone()
tick("H2D x (pageable)", lambda: dx.copy_(pageable))
b = tick("batch create", lambda: WindowBatch.from_uniform(ctx, dx, dl, labels, 50000))
st = tick("stats (5 kernels)", lambda: b.stats(0))
tick("D2H results", lambda: (hs.copy_(st[0], non_blocking=True), hc.copy_(st[1], non_blocking=True)))
tick("batch close", lambda: WindowBatch.from_uniform(ctx, dx, dl, labels, 50000).close())
t0 = time.perf_counter(); big = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True); print("pin 256MB alloc", time.perf_counter() - t0)
