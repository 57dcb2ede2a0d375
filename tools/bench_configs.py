#!/usr/bin/env python3
"""Secondary measurements (not the bench line): BASELINE configs 3, 4, 5 on one GPU, device-resident inputs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from impop_b200 import synth  # noqa: E402
from impop_b200.engine import Context, WindowBatch  # noqa: E402

ctx = Context(0)
dev = ctx.torch_device
out = {}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# config 3: 20 kb windows, 466 haplotypes (Tajima's D genome-wide); one GPU's share of 155 864 windows / 8
W3 = 19483
x, nl, pops, m, m_pad = synth.make_windows_device(ctx, 466, 20000, W3, seed=0xB203)
lab = torch.full((466,), 9, dtype=torch.uint8, device=dev)
b = WindowBatch.from_uniform(ctx, x, nl, lab, 20000)
ms = timed(lambda: b.stats(0))
ctx.timing(True); b.stats(0); pairs_ms = ctx.timing_read("pairs")[0]; prep_ms = ctx.timing_read("prep")[0]; ctx.timing(False)
out["config3"] = {"windows": W3, "nodes": int(m_pad), "ms_per_pass": ms, "pairs_kernel_ms": pairs_ms, "prep_ms": prep_ms,
                  "hap_pair_bp_per_s": W3 * 466 * 465 / 2 * 20000 / (ms * 1e-3),
                  "int8_tops": 2.0 * W3 * (466 * 467 / 2) * m_pad * 2 / (pairs_ms * 1e-3) / 1e12}
b.close(); del x, nl

# config 5: 10 000 haplotypes, 200 kb windows
n5, W5 = 10000, 4
ws = synth.make_windows(n5, 200000, 1, seed=0xB205, chunk=1)
xb = np.repeat(ws.x_bits, W5, axis=0); nlb = np.repeat(ws.node_len, W5, axis=0)
lab5 = np.full(n5, 9, dtype=np.uint8); lab5[:5000] |= 2; lab5[5000:] |= 4
b = WindowBatch.from_uniform(ctx, xb, nlb, lab5, 200000)
ms = timed(lambda: b.stats(0), reps=5, warm=2)
ctx.timing(True); b.stats(0); pairs_ms = ctx.timing_read("pairs")[0]; prep_ms = ctx.timing_read("prep")[0]; ctx.timing(False)
planes = 2
out["config5"] = {"windows": W5, "haplotypes": n5, "nodes": int(ws.m_pad), "ms_per_pass": ms, "pairs_kernel_ms": pairs_ms, "prep_ms": prep_ms,
                  "hap_pair_bp_per_s": W5 * n5 * (n5 - 1) / 2 * 200000 / (ms * 1e-3),
                  "int8_tops": 2.0 * W5 * (n5 * (n5 + 1) / 2) * ws.m_pad * planes / (pairs_ms * 1e-3) / 1e12,
                  "pair_epilogues_per_s": W5 * n5 * (n5 - 1) / 2 / (pairs_ms * 1e-3)}
b.close()

# config 4: per-site allele counts, 10^7 sites x 466 haplotypes x 5 panels
M = 10_000_000
sites_small, masks = synth.make_site_matrix(1 << 20, 466, seed=0xB204)
ds = torch.from_numpy(sites_small.view(np.int64)).to(dev).repeat((M + (1 << 20) - 1) // (1 << 20), 1)[:M].contiguous()
dm = torch.from_numpy(masks[:5].view(np.int64)).to(dev)
counts = torch.empty((M, 5), dtype=torch.int32, device=dev); freq = torch.empty((M, 5), dtype=torch.float64, device=dev)
ms = timed(lambda: ctx.site_counts(ds, dm, out_counts=counts, out_freq=freq))
bytes_alg = M * (64 + 5 * 4 + 5 * 8)
out["config4"] = {"sites": M, "ms": ms, "sites_per_s": M / (ms * 1e-3), "algorithmic_GBps": bytes_alg / (ms * 1e-3) / 1e9,
                  "hbm_frac_of_measured_6547": bytes_alg / (ms * 1e-3) / 1e9 / 6547.2}
ctx.check()
print(json.dumps(out, indent=1))
