#!/usr/bin/env python3
"""Replicates bench.py's e2e step with per-phase wall-clock timing, several consecutive steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from impop_b200 import synth  # noqa: E402
from impop_b200.engine import Context, WindowBatch, NSTATS, NCOUNTS  # noqa: E402

W = 4854
ctx = Context(0)
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 50000, W, seed=0xB201)
lab = np.full(466, 9, dtype=np.uint8); lab[pops == 0] |= 2; lab[pops == 2] |= 4
labels = torch.from_numpy(lab).to(ctx.torch_device)
hx = torch.empty(x.shape, dtype=torch.int32, pin_memory=True); hx.copy_(x)
hl = torch.empty(nl.shape, dtype=torch.int32, pin_memory=True); hl.copy_(nl)
hlab = torch.from_numpy(lab).pin_memory()
hs = torch.empty((W, NSTATS), dtype=torch.float64, pin_memory=True)
hc = torch.empty((W, NCOUNTS), dtype=torch.int64, pin_memory=True)
dx, dl, dlab = torch.empty_like(x), torch.empty_like(nl), torch.empty_like(labels)
torch.cuda.synchronize()
for algo in (0, 1, 0):
    for it in range(4):
        t = [time.perf_counter()]
        dx.copy_(hx, non_blocking=True); dl.copy_(hl, non_blocking=True); dlab.copy_(hlab, non_blocking=True)
        t.append(time.perf_counter())
        b = WindowBatch.from_uniform(ctx, dx, dl, dlab, 50000)
        t.append(time.perf_counter())
        s, c = b.stats(algo)
        t.append(time.perf_counter())
        hs.copy_(s, non_blocking=True); hc.copy_(c, non_blocking=True)
        t.append(time.perf_counter())
        torch.cuda.current_stream().synchronize()
        t.append(time.perf_counter())
        b.close()
        t.append(time.perf_counter())
        d = [(t[i + 1] - t[i]) * 1e3 for i in range(len(t) - 1)]
        print(f"algo {algo} step {it}: enqueue-h2d {d[0]:.2f} create {d[1]:.2f} stats-call {d[2]:.2f} d2h-enq {d[3]:.2f} sync {d[4]:.2f} close {d[5]:.2f} total {sum(d):.2f} ms")
