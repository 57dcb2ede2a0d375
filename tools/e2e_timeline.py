#!/usr/bin/env python3
"""GPU timeline of one pipelined end-to-end step (events per sub-batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from impop_b200 import synth
from impop_b200.engine import Context, WindowBatch, NSTATS, NCOUNTS
W = 4854; nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
wts = np.ones(nsub)
if nsub >= 4 and "--plain" not in sys.argv:
    wts[-2], wts[-1] = 0.6, 0.3                                  # as bench.py: small last sub-batches, short tail
cuts = [0] + [int(v) for v in np.round(np.cumsum(wts) / wts.sum() * W)]; cuts[-1] = W
dlabs = [torch.empty_like(labels) for _ in range(nsub)]
def step(record=False):
    live = []; evs = []; host = []
    t00 = time.perf_counter()
    for k in range(nsub):
        lo, hi = cuts[k], cuts[k + 1]; st = streams[k % 2]
        with torch.cuda.stream(st):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            dlabs[k].copy_(hlab, non_blocking=True); dl[lo:hi].copy_(hl[lo:hi], non_blocking=True); dx[lo:hi].copy_(hx[lo:hi], non_blocking=True)
            e[1].record()
            h0 = time.perf_counter()
            b = WindowBatch.from_uniform(ctx, dx[lo:hi], dl[lo:hi], dlabs[k], 50000, node_len_host=hl[lo:hi], stream=st)
            h1 = time.perf_counter()
            b.stats(0, stream=st, out_stats=ds[lo:hi], out_counts=dc[lo:hi])
            e[2].record()
            hs[lo:hi].copy_(ds[lo:hi], non_blocking=True); hc[lo:hi].copy_(dc[lo:hi], non_blocking=True)
            e[3].record()
            h2 = time.perf_counter()
        live.append(b); evs.append(e); host.append((h0 - t00, h1 - t00, h2 - t00))
    for st in streams: st.synchronize()
    for b in live: b.close()
    return evs, host
for _ in range(3): step()
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record(); torch.cuda.synchronize()
ctx.timing(True)
evs, host = step(True)
for kname in ('prep', 'pairs', 'sums', 'finalize'):
    ms, cnt = ctx.timing_read(kname)
    print(f'kernel {kname}: {cnt} launches, total {ms:.3f} ms')
ctx.timing(False)
for k, (e, h) in enumerate(zip(evs, host)):
    t = [base.elapsed_time(x_) for x_ in e]
    print(f"sub {k} stream {k % 2}: h2d {t[0]:7.3f} -> {t[1]:7.3f} | kernels end {t[2]:7.3f} | d2h end {t[3]:7.3f}   || host: create {h[0] * 1e3:6.3f}->{h[1] * 1e3:6.3f}, enqueued {h[2] * 1e3:6.3f} ms")
