#!/usr/bin/env python3
"""Short resident-only timing of the fused path on chr2-shaped windows, original vs compacted columns."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from impop_b200 import synth, ingest
from impop_b200.engine import Context, WindowBatch
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4854
ctx = Context(0); dev = ctx.torch_device
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 50000, W, seed=0xB201)
lab = np.full(466, 9, dtype=np.uint8); lab[pops == 0] |= 2; lab[pops == 2] |= 4
labels = torch.from_numpy(lab).to(dev)
def run(xd, ld, tag, **kw):
    b = WindowBatch.from_uniform(ctx, xd, ld, labels, 50000, **kw)
    for _ in range(3): st, ct = b.stats(0)
    ctx.check(); ctx.timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): st, ct = b.stats(0)
    e1.record(); torch.cuda.synchronize()
    per = {k: ctx.timing_read(k)[0] / 10 for k in ("prep", "pairs", "sums", "finalize")}
    ctx.timing(False)
    print(tag, "ms_per_step %.3f" % (e0.elapsed_time(e1) / 10), {k: round(v, 3) for k, v in per.items()}, "nodes", ld.shape[1], flush=True)
    b.close()
    return st.cpu().numpy(), ct.cpu().numpy()
s0, c0 = run(x, nl, "original ")
xh = x.cpu().numpy().view(np.uint32); lh = nl.cpu().numpy().view(np.uint32)
PLAIN = "--plain" in sys.argv
t0 = time.perf_counter(); cu = ingest.compact_uniform(xh, lh, pairs=not PLAIN); t1 = time.perf_counter()
xc, lc, mo = cu.x, cu.node_len, cu.m
KW = cu.batch_kwargs(upload=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev))
print("compaction %.3f s on %d threads; nodes %d -> max %d (pitch %d words)" % (t1 - t0, len(os.sched_getaffinity(0)), m, mo.max(), xc.shape[2]), flush=True)
xd = torch.from_numpy(xc.view(np.int32)).to(dev); ld = torch.from_numpy(lc.view(np.int32)).to(dev)
s1, c1 = run(xd, ld, "compacted", **KW)
s0, s1 = s0[:, :19], s1[:, :19]            # column 19 (variant sites) depends on the node order
print("counts equal:", bool((c0 == c1).all()), " stats max rel diff:", float(np.nanmax(np.abs(s0 - s1) / np.maximum(np.abs(s0), 1e-300))),
      " nan pattern equal:", bool((np.isnan(s0) == np.isnan(s1)).all()))
