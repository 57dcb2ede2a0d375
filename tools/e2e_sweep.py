#!/usr/bin/env python3
"""End-to-end step time (pinned host buffers -> results on the host) against the number of sub-batches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from impop_b200 import synth
from impop_b200.engine import Context, WindowBatch, NSTATS, NCOUNTS
W = 4854
ctx = Context(0); dev = ctx.torch_device
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 50000, W, seed=0xB201)
lab = np.full(466, 9, dtype=np.uint8); lab[pops == 0] |= 2; lab[pops == 2] |= 4
labels = torch.from_numpy(lab).to(dev)
hx = torch.empty(x.shape, dtype=torch.int32, pin_memory=True); hx.copy_(x)
hl = torch.empty(nl.shape, dtype=torch.int32, pin_memory=True); hl.copy_(nl)
hlab = torch.from_numpy(lab).pin_memory()
hs = torch.empty((W, NSTATS), dtype=torch.float64, pin_memory=True); hc = torch.empty((W, NCOUNTS), dtype=torch.int64, pin_memory=True)
dx, dl = torch.empty_like(x), torch.empty_like(nl)
ds = torch.empty((W, NSTATS), dtype=torch.float64, device=dev); dc = torch.empty((W, NCOUNTS), dtype=torch.int64, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
for nsub in (4, 6, 8, 10, 12, 16, 24):
    cuts = [int(v) for v in np.linspace(0, W, nsub + 1)]
    dlabs = [torch.empty_like(labels) for _ in range(nsub)]
    def step(host_only=False):
        live = []; t_host = 0.0
        for k in range(nsub):
            lo, hi = cuts[k], cuts[k + 1]
            st = streams[k % 2]
            with torch.cuda.stream(st):
                dx[lo:hi].copy_(hx[lo:hi], non_blocking=True); dl[lo:hi].copy_(hl[lo:hi], non_blocking=True); dlabs[k].copy_(hlab, non_blocking=True)
                t0 = time.perf_counter()
                b = WindowBatch.from_uniform(ctx, dx[lo:hi], dl[lo:hi], dlabs[k], 50000, node_len_host=hl[lo:hi], stream=st)
                t_host += time.perf_counter() - t0
                b.stats(0, stream=st, out_stats=ds[lo:hi], out_counts=dc[lo:hi])
                hs[lo:hi].copy_(ds[lo:hi], non_blocking=True); hc[lo:hi].copy_(dc[lo:hi], non_blocking=True)
            live.append(b)
        t1 = time.perf_counter()
        for st in streams: st.synchronize()
        t_sync = time.perf_counter() - t1
        for b in live: b.close()
        return t_host, t_sync
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter(); th = ts = 0.0
    for _ in range(8):
        a, b_ = step(); th += a; ts += b_
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    print(f"sub-batches {nsub:3d}: e2e {dt * 1e3:7.3f} ms/step   host time in batch_create {th / 8 * 1e3:6.3f} ms   final sync wait {ts / 8 * 1e3:6.3f} ms")
