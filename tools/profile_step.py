#!/usr/bin/env python3
"""Short run of the fused window path for ncu: `--windows W` HPRC-shaped windows, `--reps R` passes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from impop_b200 import synth  # noqa: E402
from impop_b200.engine import Context, WindowBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=592)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--algo", type=int, default=0)
ap.add_argument("--n", type=int, default=466)
ap.add_argument("--length", type=int, default=50000)
ap.add_argument("--compact", action="store_true", help="compact the columns at ingest (impop_compact_scan / _fill)")
ap.add_argument("--plain", action="store_true", help="with --compact: the plain form (constant columns merged only)")
ap.add_argument("--full-pitch", action="store_true", help="with --compact: every window uses all columns of the shared pitch (rows contiguous: prep_rows' block path)")
args = ap.parse_args()
ctx = Context(0)
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, args.n, args.length, args.windows, seed=0xB201)
KW = {}
if args.compact:
    from impop_b200 import ingest  # noqa: E402
    cu = ingest.compact_uniform(x.cpu().numpy().view(np.uint32), nl.cpu().numpy().view(np.uint32), pairs=not args.plain)
    x = torch.from_numpy(cu.x.view(np.int32)).to(ctx.torch_device)
    nl = torch.from_numpy(cu.node_len.view(np.int32)).to(ctx.torch_device)
    KW = cu.batch_kwargs(upload=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.torch_device))
    if args.full_pitch:
        KW["m"] = None
    print("compacted: nodes", m, "->", int(cu.m.max()), flush=True)
lab = np.full(args.n, 9, dtype=np.uint8)
lab[pops == 0] |= 2
lab[pops == 2] |= 4
batch = WindowBatch.from_uniform(ctx, x, nl, torch.from_numpy(lab).to(ctx.torch_device), args.length, **KW)
for _ in range(args.reps):
    stats, counts = batch.stats(args.algo)
ctx.check()
print("ok", float(stats[0, 0]), int(counts[0, 7]), "items", batch.items)
