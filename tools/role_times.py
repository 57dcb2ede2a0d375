#!/usr/bin/env python3
"""Role-time breakdown of the pairs kernel (needs a -DIMPOP_PROFILE_ROLES build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from impop_b200 import synth
from impop_b200.engine import Context, WindowBatch
W = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4854
ctx = Context(0)
KW = {}
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 50000, W, seed=0xB201)
if "--compact" in sys.argv:            # columns compacted at ingest (impop_compact_scan / _fill)
    from impop_b200 import ingest
    cu = ingest.compact_uniform(x.cpu().numpy().view(np.uint32), nl.cpu().numpy().view(np.uint32), pairs="--plain" not in sys.argv)
    x = torch.from_numpy(cu.x.view(np.int32)).to(ctx.torch_device); nl = torch.from_numpy(cu.node_len.view(np.int32)).to(ctx.torch_device)
    KW = cu.batch_kwargs(upload=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.torch_device))
    print("compacted: nodes", m, "->", int(cu.m.max()))
lab = np.full(466, 9, dtype=np.uint8); lab[pops == 0] |= 2; lab[pops == 2] |= 4
b = WindowBatch.from_uniform(ctx, x, nl, torch.from_numpy(lab).to(ctx.torch_device), 50000, **KW)
for _ in range(3):
    b.stats(0)
ctx.check()
out = np.zeros((148, 16), dtype=np.int64)
ctx._call("impop_debug_role_times", C.c_void_p(out.ctypes.data), 148)
names = ["producer A (warp 0)", "producer B (warp 4)", "MMA issuer", "epilogue team 0", "epilogue team 1"]
labels = {0: ("wait empty", "expand+arrive", "item setup"), 1: ("wait empty", "expand+arrive", "item setup"),
          2: ("wait full", "issue", "wait acc_empty"), 3: ("wait acc_full", "chunks", "setup+table+reduce"), 4: ("wait acc_full", "chunks", "setup+table+reduce")}
items = b.items / 148
for k, nm in enumerate(names):
    v = out[:, 3 * k: 3 * k + 3].mean(axis=0)
    tot = v.sum()
    print(f"{nm:22s} total {tot / 1e3:9.1f} kcyc  per item {tot / items:8.0f} | " + ", ".join(f"{l} {x_ / items:7.0f}" for l, x_ in zip(labels[k], v)))
