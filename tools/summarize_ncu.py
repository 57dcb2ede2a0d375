#!/usr/bin/env python3
"""Text summary of one kernel capture in an .ncu-rep (run where ncu is installed, no GPU needed):
key raw metrics + executed-instruction / stall-sample shares per SASS region and per opcode.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/r1_xxx.txt
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
block = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# ncu summary of {rep}")
units = rows[1] if len(rows) > 2 else [""] * len(hdr)
for key in want:
    for h, u, v in zip(hdr, units, vals):
        if h == key:
            print(f"{h} = {v} {u if u != v else ''}".rstrip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


def opcode(r):
    t = r[ix["Source"]].split()
    if not t:
        return "?"
    return t[1] if t[0].startswith("@") and len(t) > 1 else t[0]


tot_s = sum(f(r, "# Samples") for r in data) or 1.0
tot_i = sum(f(r, "Instructions Executed") for r in data) or 1.0
print(f"\n# source page: {len(data)} SASS instructions, {int(tot_i)} warp-level instructions executed, {int(tot_s)} stall samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(f(r, h) for r in data) for h in stalls}
print("stall samples by reason: " + ", ".join(f"{h[6:]}={int(v)}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
print(f"\n# regions of {block} SASS instructions (address order): share of samples / executed instructions, top stalls, marker opcodes")
for b in range(0, len(data), block):
    blk = data[b:b + block]
    s = sum(f(r, "# Samples") for r in blk)
    i = sum(f(r, "Instructions Executed") for r in blk)
    if s / tot_s < 0.005 and i / tot_i < 0.005:
        continue
    st = collections.Counter({h[6:]: sum(f(r, h) for r in blk) for h in stalls})
    marks = sorted({opcode(r) for r in blk if any(k in opcode(r) for k in
                    ("MUFU", "UTC", "LDTM", "BAR", "SYNCS", "STS", "LDG", "SHFL", "ATOM", "DFMA", "STG", "LDS", "MEMBAR", "VOTE", "NANOSLEEP"))})
    print(f"{b:5d}: samples {100 * s / tot_s:5.1f}%  instr {100 * i / tot_i:5.1f}%  | "
          + ", ".join(f"{k}:{int(v)}" for k, v in st.most_common(3)) + " | " + " ".join(marks)[:100])
by = collections.defaultdict(lambda: [0.0, 0.0])
for r in data:
    by[opcode(r)][0] += f(r, "# Samples")
    by[opcode(r)][1] += f(r, "Instructions Executed")
print("\n# top opcodes by executed instructions")
for op, (s, i) in sorted(by.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{op:32s} instr {100 * i / tot_i:5.1f}%   samples {100 * s / tot_s:5.1f}%")
