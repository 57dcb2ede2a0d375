#!/usr/bin/env python3
"""The real file-to-table path, stage by stage: window graphs as GFA text on disk -> native reader (all host cores) ->
ingest (column compaction) -> flat container -> (re-load, memory-mapped) -> upload -> fused kernels -> the wrappers' TSVs.
Replaces the per-window loops of run_h-fst.sh:155-194 / run_tajd.sh:103-198 (3-6 process spawns and one text table per
BED row).  The window set is `--windows` chr2-shaped windows (466 haplotypes, 50 kb); the GFA stage is extrapolated to a
whole chromosome from the measured per-window cost, the container stages are measured at full size with `--full`.

    python tools/bench_pipeline.py [--windows 256] [--full 4854] [--out profiles/r2_pipeline.json]
"""
import argparse, json, os, subprocess, sys, tempfile, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from impop_b200 import ingest, synth, windows

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=256)
ap.add_argument("--full", type=int, default=4854)
ap.add_argument("--out", default="")
args = ap.parse_args()
cores = len(os.sched_getaffinity(0))
N, L = 466, 50_000
res = {"cores": cores, "haplotypes": N, "window_bp": L}

with tempfile.TemporaryDirectory() as tmp:
    # ---- stage A: GFA text on disk -> TSVs through the command line (what a user runs)
    W = args.windows
    ws = synth.make_windows(N, L, min(W, 32), seed=0xB200 + 1)
    names = synth.haplotype_names(N, "chr2", 0, L)
    asm = synth.assembly_names(range(N))
    pops, _ = synth.panel(N)
    open(os.path.join(tmp, "a.txt"), "w").write("\n".join(a for a, p in zip(asm, pops) if p == 0) + "\n")
    open(os.path.join(tmp, "b.txt"), "w").write("\n".join(a for a, p in zip(asm, pops) if p == 2) + "\n")
    listing, gfa_bytes = [], 0
    t0 = time.perf_counter()
    for w in range(W):
        path = os.path.join(tmp, f"w{w}.gfa")
        src = w % ws.windows
        if w < ws.windows:
            with open(path, "w") as fh:
                ingest.write_gfa(fh, names, ws.dense(src)[:, :ws.m], ws.node_len[src, :ws.m])
        else:
            os.link(os.path.join(tmp, f"w{src}.gfa"), path)         # same text again: the reader does not know
        gfa_bytes += os.path.getsize(path)
        listing.append(f"{windows.region_name('chr2', w * L, (w + 1) * L)}\t{path}")
    open(os.path.join(tmp, "windows.tsv"), "w").write("\n".join(listing) + "\n")
    res["gfa"] = {"windows": W, "text_bytes": gfa_bytes, "write_s": time.perf_counter() - t0}
    cmd = [sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "impop-windows.py"),
           "--gfa-list", os.path.join(tmp, "windows.tsv"), "-a", os.path.join(tmp, "a.txt"), "-b", os.path.join(tmp, "b.txt"),
           "--fst-out", os.path.join(tmp, "fst.tsv"), "--tajd-out", os.path.join(tmp, "tajd.tsv"), "--pi-out", os.path.join(tmp, "pi.tsv"),
           "--save-batch", os.path.join(tmp, "batch.impw"), "--timings"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    res["gfa"]["command_wall_s"] = time.perf_counter() - t0
    res["gfa"]["stages"] = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""
    res["gfa"]["rc"] = r.returncode
    res["gfa"]["parse_ms_per_window_per_core"] = None
    for part in res["gfa"]["stages"].split(","):
        if " parse " in part or part.strip().startswith("impop-windows timings (s): parse"):
            try:
                sec = float(part.strip().split()[-1])
                res["gfa"]["parse_s"] = sec
                res["gfa"]["parse_ms_per_window_per_core"] = sec * 1e3 * min(cores, W) / W
            except ValueError:
                pass
    res["gfa"]["chr2_extrapolation"] = ("parse of 4854 windows at the measured per-window cost: %.1f s on %d cores"
                                        % ((res["gfa"].get("parse_s") or 0.0) * 4854 / W, cores))
    # ---- stage B: the container path at full size (what every later run of the same windows pays)
    F = args.full
    big = synth.make_windows(N, L, 64, seed=0xB200 + 5)
    graphs = []
    for w in range(F):
        src = w % 64
        g = ingest.GraphWindow([s.replace(":0-50000", f":{w * L}-{(w + 1) * L}") for s in names] if w < 2 else None,
                               big.x_bits[src], big.node_len[src, :big.m], None, windows.region_name("chr2", w * L, (w + 1) * L), L)
        graphs.append(g)
    base_names = [s.split(":")[0] for s in names]
    for w, g in enumerate(graphs):
        if g.names is None:
            g.names = [f"{b}:{w * L}-{(w + 1) * L}" for b in base_names]
    t0 = time.perf_counter()
    xb = np.stack([g.x_bits for g in graphs]); nl = np.zeros((F, big.m_pad), np.uint32); nl[:, :big.m] = np.stack([g.node_len for g in graphs])
    cu = ingest.compact_uniform(xb, nl)                      # affine form: bubbles merged, constants in C
    t1 = time.perf_counter()
    cg = [ingest.GraphWindow(g.names, cu.x[w], cu.node_len[w, :int(cu.m[w])], None, g.region, L, int(cu.site_runs[w]),
                             cu.row_adj[w], int(cu.win_const[w]), cu.col_mult[w, :int(cu.m[w])]) for w, g in enumerate(graphs)]
    flat_path = os.path.join(tmp, "chr2.impw")
    t2 = time.perf_counter()
    ingest.save_flat(flat_path, cg)
    t3 = time.perf_counter()
    res["container"] = {"windows": F, "compaction_s": t1 - t0, "save_flat_s": t3 - t2, "file_bytes": os.path.getsize(flat_path)}
    cmd = [sys.executable, cmd[1], "--batch", flat_path, "-a", os.path.join(tmp, "a.txt"), "-b", os.path.join(tmp, "b.txt"),
           "--fst-out", os.path.join(tmp, "fst2.tsv"), "--tajd-out", os.path.join(tmp, "tajd2.tsv"), "--pi-out", os.path.join(tmp, "pi2.tsv"), "--timings"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    res["container"]["command_wall_s"] = time.perf_counter() - t0
    res["container"]["stages"] = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else r.stderr[-300:]
    res["container"]["rc"] = r.returncode
    t0 = time.perf_counter()
    fb = ingest.load_flat(flat_path)
    lab = fb.labels(pop_a=["S00000#"], pop_b=["S00001#"])
    res["container"]["load_flat_plus_labels_s"] = time.perf_counter() - t0
    res["container"]["rows_written"] = sum(1 for _ in open(os.path.join(tmp, "fst2.tsv"))) - 1 if r.returncode == 0 else 0
print(json.dumps(res, indent=1))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
