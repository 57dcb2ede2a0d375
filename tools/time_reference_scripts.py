#!/usr/bin/env python3
"""BASELINE.md section 3, steps 2-3, in the BUILD container (the reference tree does not travel to the GPU box): the
unmodified reference scripts timed on oracle-written similarity tables of synthetic HPRC-shaped windows -- one process
per window as the wrappers run them (run_pica2_impg.sh:175, run_h-fst.sh:74-85), and fanned out over windows with a
process pool.  Writes profiles/r2_reference_scripts.json.

    python tools/time_reference_scripts.py [--reference /root/reference] [--windows 8]
"""
import argparse, json, os, subprocess, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from impop_b200 import synth
from oracle import similarity

ap = argparse.ArgumentParser()
ap.add_argument("--reference", default="/root/reference")
ap.add_argument("--windows", type=int, default=8)
args = ap.parse_args()
ref = os.path.join(args.reference, "scripts")
cores = len(os.sched_getaffinity(0))
out = {"where": "build container (no GPU)", "cores": cores, "python": sys.version.split()[0], "cases": {}}


def timed(cmd, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r


with tempfile.TemporaryDirectory() as tmp:
    for n, L, tag in ((90, 100_000, "config1_n90_100kb"), (466, 50_000, "config2_n466_50kb")):
        ws = synth.make_windows(n, L, args.windows, seed=0xB200 + (0 if n == 90 else 1))
        names = synth.haplotype_names(n, "chr2", 0, L)
        tsvs = []
        for w in range(args.windows):
            res = similarity.pairwise(ws.dense(w), ws.node_len[w])
            p = os.path.join(tmp, f"{tag}.{w}.sim.tsv")
            similarity.write_similarity_tsv(p, names, res)
            tsvs.append(p)
        pops, _ = synth.panel(n)
        asm = synth.assembly_names(range(n))
        fa, fb = os.path.join(tmp, f"{tag}.a.txt"), os.path.join(tmp, f"{tag}.b.txt")
        open(fa, "w").write("\n".join(a for a, p in zip(asm, pops) if p == 0) + "\n")
        open(fb, "w").write("\n".join(a for a, p in zip(asm, pops) if p == 2) + "\n")
        units = n * (n - 1) / 2 * L
        case = {"haplotypes": n, "window_bp": L, "rows_per_table": n * (n - 1) // 2, "table_bytes": os.path.getsize(tsvs[0]), "scripts": {}}
        cmds = {
            "pica2.py -t 1.0 -l L": lambda t: [sys.executable, os.path.join(ref, "pica2.py"), t, "-t", "1.0", "-l", str(L), "-d", tmp],
            "pica2.py -t 0.999 -r 5 -l L": lambda t: [sys.executable, os.path.join(ref, "pica2.py"), t, "-t", "0.999", "-r", "5", "-l", str(L), "-d", tmp],
            "h-fst.py -a AFR -b EAS -l L": lambda t: [sys.executable, os.path.join(ref, "h-fst.py"), t, "-a", fa, "-b", fb, "-l", str(L), "-d", tmp],
            "af.py --threshold 0.9995": lambda t: [sys.executable, os.path.join(ref, "af.py"), "--input", t, "--threshold", "0.9995", "--output", os.path.join(tmp, "af.out")],
        }
        for label, mk in cmds.items():
            single, r = timed(mk(tsvs[0]))
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=cores) as pool:          # one reference process per window, `cores` at a time
                list(pool.map(lambda t: subprocess.run(mk(t), capture_output=True), tsvs))
            fan = time.perf_counter() - t0
            case["scripts"][label] = {"single_process_wall_s": single, "hap_pair_bp_per_s_single": units / single,
                                      "pool_windows": len(tsvs), "pool_wall_s": fan, "hap_pair_bp_per_s_pool": units * len(tsvs) / fan,
                                      "rc": r.returncode, "stdout": r.stdout.strip().splitlines()[-1][:60] if r.stdout.strip() else ""}
        out["cases"][tag] = case
    t, r = timed([sys.executable, os.path.join(ref, "tj_d.py"), "-n", "446", "-p", "0.59146123", "-S", "20"])
    out["tj_d.py -n 446 -p 0.59146123 -S 20"] = {"single_process_wall_s": t, "stdout": r.stdout.strip()}
out["note"] = ("Python reduction step only: the external similarity step (odgi / impg, not installed, not in the reference tree) is not in "
               "these times.  Same tables, same flags as bench.py's config 1 / 2 workloads.")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_reference_scripts.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
