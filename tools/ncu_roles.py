#!/usr/bin/env python3
"""Executed-instruction mix of the pairs kernel per warp role from an .ncu-rep (source page), normalised per window and SM
and, for the epilogue, per 32-pair warp operation.  Wait loops are excluded (their counts come from an instrumented replay).

    python tools/ncu_roles.py gpurun_out/prof.ncu-rep <windows in the launch> [pair-ops per window per SM, default 3840]
"""
import collections, csv, io, subprocess, sys

rep, W = sys.argv[1], float(sys.argv[2])
PAIROPS = float(sys.argv[3]) if len(sys.argv) > 3 else 3840.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def ie(r):
    try:
        return float(r[ix["Instructions Executed"]])
    except Exception:
        return 0.0


def op(r):
    t = r[ix["Source"]].split()
    return (t[1] if t[0].startswith("@") else t[0]) if t else "?"


n = len(data)
poll = [False] * n
for k, r in enumerate(data):
    if "NANOSLEEP" in r[ix["Source"]]:
        for j in range(max(0, k - 4), min(n, k + 23)):
            poll[j] = True
marks = [k for k, r in enumerate(data) if "USETMAXREG" in r[ix["Source"]]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
for key in ("gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    for h, v in zip(rr[0], rr[-1]):
        if h == key:
            print(f"{h} = {v}")
SM = 148.0
FP64 = ("DFMA", "DADD", "DMUL")
tot_cycles = 0.0
for a, b, name, norm, label in ((0, marks[0], "prologue", 1.0, "per window per SM"), (marks[0], marks[1], "producers", 1.0, "per window per SM"),
                                (marks[1], marks[2], "MMA / table / loaders", 1.0, "per window per SM"),
                                (marks[2], n, "epilogue", PAIROPS, "per 32-pair warp operation")):
    c = collections.Counter()
    for k in range(a, b):
        if not poll[k]:
            c[op(data[k])] += ie(data[k])
    tot = sum(c.values()) / W
    f64 = sum(v for o, v in c.items() if o.startswith(FP64)) / W
    issue = tot + f64                                 # an fp64 instruction holds the issue port for two cycles
    tot_cycles += issue
    print(f"\n{name}: {tot:9.0f} instructions per window per SM ({f64:.0f} fp64) = {issue:9.0f} issue cycles")
    print(f"   {label}: " + ", ".join("%s %.2f" % (o, v / W / norm) for o, v in c.most_common(24)))
print(f"\nissue cycles per window and SM sub-partition (sum / 4): {tot_cycles / 4:.0f}")
