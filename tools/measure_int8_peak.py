#!/usr/bin/env python3
"""Measure the dense int8 tensor peak the way MEASURED_PEAKS.json's bf16 entry was measured
(SURVEY.md 8 d): torch._int_mm 8192^3, best of 10 (burst) and back to back for ~4 s (sustained).
Writes profiles/int8_peak.json (the roofline denominator for the int8 pairwise kernel)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 8192
a = torch.randint(-8, 8, (N, N), dtype=torch.int8, device="cuda")
b = torch.randint(-8, 8, (N, N), dtype=torch.int8, device="cuda")
for _ in range(3):
    torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
ops = 2.0 * N ** 3
burst = ops / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0, reps = time.time(), 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(50):
        torch._int_mm(a, b)
    reps += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sustained = ops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"int8_tops": burst, "int8_tops_sustained": sustained, "how": "torch._int_mm 8192^3 (2*N^3), best of 10 and 4 s loop",
       "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__}
dst = os.path.join(ROOT, "gpurun_out", "int8_peak.json") if "--scratch" in sys.argv else os.path.join(ROOT, "profiles", "int8_peak.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
