#!/usr/bin/env python3
"""Regenerate tests/golden/*.json by running the UNMODIFIED reference scripts.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Everything numeric is stored as float.hex() so the fixtures are bit-exact.  Inputs
(the small similarity tables / window matrices) are stored next to the outputs, so
the GPU-box tests never need the reference tree.
"""
from __future__ import annotations

import io
import json
import os
import re
import sys
import tempfile
from contextlib import redirect_stderr, redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload, similarity  # noqa: E402
from impop_b200 import synth  # noqa: E402


def hx(v):
    if isinstance(v, (list, tuple)):
        return [hx(u) for u in v]
    if isinstance(v, dict):
        return {k: hx(u) for k, u in v.items()}
    if isinstance(v, (float, np.floating)):
        return float(v).hex()
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def extract_f6():
    """The 6-sequence table and population lists embedded in hudson/example_fst_methods.py:7-37."""
    src = open(os.path.join(refload.REFERENCE_SCRIPTS, "hudson", "example_fst_methods.py")).read()
    blocks = re.findall(r"cat > (\S+) << 'EOF'\n(.*?)\nEOF", src, flags=re.S)
    out = {name: body for name, body in blocks}
    return out["example_similarities.tsv"], out["pop_A.txt"].split(), out["pop_B.txt"].split()


def run_cli(mod, argv):
    """Run a reference script's __main__ body / main() and capture (stdout, stderr, exit code)."""
    import runpy
    so, se = io.StringIO(), io.StringIO()
    code = 0
    old = sys.argv
    sys.argv = argv
    try:
        with redirect_stdout(so), redirect_stderr(se):
            try:
                runpy.run_path(mod.__file__, run_name="__main__")
            except SystemExit as exc:
                code = exc.code if isinstance(exc.code, int) else (0 if exc.code is None else 1)
    finally:
        sys.argv = old
    return so.getvalue(), se.getvalue(), code


def grouping_is_transitive(path, thr, r):
    """pica2's greedy grouping (pica2.py:94-112) pops set elements in hash order; its result is only
    defined when `sim > thr` is an equivalence relation on the table."""
    from oracle import popstats
    names, mat, _ = popstats.parse_similarity_tsv(path)
    mat = popstats.py_round_matrix(mat, r)
    adj = np.nan_to_num(mat, nan=-np.inf) > thr
    np.fill_diagonal(adj, True)
    reach = adj.copy()
    for _ in range(len(names)):
        reach = reach | ((reach.astype(np.int64) @ reach.astype(np.int64)) > 0)
    return bool((reach == adj).all())


def table_cases(tsv_text, pop_a_ids, pop_b_ids, tmp, tag, thresholds, expand=False, lengths=(None, 50000),
                rounds=(None, 3), af_thresholds=(0.9995, 0.9994, 1.0, 0.99)):
    pica2, hfst, af, hud = (refload.load(k) for k in ("pica2", "hfst", "af", "hud"))
    path = os.path.join(tmp, f"{tag}.tsv")
    with open(path, "w") as fh:
        fh.write(tsv_text if tsv_text.endswith("\n") else tsv_text + "\n")
    case = {"tsv": open(path).read(), "pop_a": list(pop_a_ids), "pop_b": list(pop_b_ids), "expand": expand}
    # --- pica2
    case["pica2"] = []
    for thr in thresholds:
        for L in lengths:
            for r in rounds:
                with redirect_stdout(io.StringIO()):
                    sim, el, cnt = pica2.read_similarity_file(path)
                    try:
                        pi, pps = pica2.analyze_similarity_matrix(sim, el, cnt, threshold=thr, sequence_length=L,
                                                                  round_digits=r)
                    except ZeroDivisionError:
                        pi, pps = "ZeroDivisionError", None
                case["pica2"].append({"threshold": thr, "L": L, "round": r, "pi": hx(pi), "pi_per_site": hx(pps),
                                      "transitive": grouping_is_transitive(path, thr, r)})
    # --- h-fst (function level)
    case["hfst"] = []
    with redirect_stderr(io.StringIO()):
        sims, seqs = hfst.read_similarity_file(path)
        if expand:
            pa, miss_a = hfst.expand_population(set(pop_a_ids), seqs)
            pb, miss_b = hfst.expand_population(set(pop_b_ids), seqs)
            case["expanded"] = {"a": sorted(pa), "b": sorted(pb), "missing_a": sorted(miss_a), "missing_b": sorted(miss_b)}
        else:
            pa, pb = set(pop_a_ids), set(pop_b_ids)
        for L in lengths:
            for r in rounds:
                res = hfst.calculate_fst(sims, set(pa), set(pb), sequence_length=L, round_digits=r)
                case["hfst"].append({"L": L, "round": r, "res": hx(res)})
        d_within = hfst.calculate_diversity(sims, set(pa))
        d_between = hfst.calculate_diversity(sims, set(pa), set(pb))
        case["diversity"] = {"within_a": hx(list(d_within)), "between": hx(list(d_between))}
        hd = hud.calculate_fst(sims, set(pa), set(pb), sequence_length=None, method="direct") \
            if "method" in hud.calculate_fst.__code__.co_varnames else None
        case["hud_direct"] = hx({k: v for k, v in hd.items() if isinstance(v, float)}) if isinstance(hd, dict) else None
    # --- af
    case["af"] = []
    rows, samples = af.load_pairs(path)
    for thr in af_thresholds:
        clusters = af.cluster(rows, samples, thr)
        summary = af.build_summary(clusters)
        case["af"].append({"threshold": thr, "summary": [[cid, cnt, hx(fr), mem] for cid, cnt, fr, mem in summary]})
    return case


def cli_cases(tmp, f6_tsv, pop_a, pop_b):
    """stdout / exit code of the four CLIs on the F6 fixture."""
    pica2, hfst, tj, af, hud = (refload.load(k) for k in ("pica2", "hfst", "tj_d", "af", "hud"))
    tsv = os.path.join(tmp, "f6cli.tsv")
    open(tsv, "w").write(f6_tsv + "\n")
    fa, fb = os.path.join(tmp, "pa.txt"), os.path.join(tmp, "pb.txt")
    open(fa, "w").write("\n".join(pop_a) + "\n")
    open(fb, "w").write("\n".join(pop_b) + "\n")
    logd = os.path.join(tmp, "logs")
    out = {}
    runs = {
        "pica2_t0999": (pica2, ["pica2.py", tsv, "-t", "0.999", "-d", logd]),
        "pica2_t0999_l_r5": (pica2, ["pica2.py", tsv, "-t", "0.999", "-l", "1000000", "-r", "5", "-d", logd]),
        "pica2_t1_l": (pica2, ["pica2.py", tsv, "-t", "1.0", "-l", "100000", "-d", logd]),
        "pica2_missing_file": (pica2, ["pica2.py", os.path.join(tmp, "nope.tsv"), "-d", logd]),
        "hfst_f6": (hfst, ["h-fst.py", tsv, "-a", fa, "-b", fb, "-d", logd]),
        "hud_direct": (hud, ["hud.py", tsv, "-a", fa, "-b", fb, "-l", "1000000", "-m", "direct", "-d", logd]),
        "tjd_doc": (tj, ["tj_d.py", "-n", "446", "-p", "0.59146123", "-S", "20", "--show-components"]),
        "tjd_s0": (tj, ["tj_d.py", "-n", "446", "-p", "0.5", "-S", "0"]),
        "af_09995": (af, ["af.py", "--input", tsv, "--threshold", "0.9995", "--details", os.path.join(tmp, "det.tsv")]),
    }
    for key, (mod, argv) in runs.items():
        so, se, code = run_cli(mod, argv)
        out[key] = {"argv": [a.replace(tmp, "$TMP") for a in argv], "stdout": so.replace(tmp, "$TMP"), "code": code}
    out["af_09995"]["details"] = open(os.path.join(tmp, "det.tsv")).read()
    return out


def tajima_grid():
    tj = refload.load("tj_d")
    rows = []
    for n in (2, 3, 10, 90, 446, 466, 1000, 10000):
        for S in (0.0, 1.0, 20.0, 134.0, 1200.0, 10000.0):
            for pi in (0.0, 1e-6, 0.5, 0.59146123, 85.0):
                d, c = tj.tajimas_d(n, S, pi, return_components=True)
                rows.append({"n": n, "S": hx(S), "pi": hx(pi), "D": hx(d),
                             "parts": hx([c.a1, c.a2, c.b1, c.b2, c.c1, c.c2, c.e1, c.e2, c.numerator, c.denominator])})
    return rows


def canonical_cases():
    hfst = refload.load("hfst")
    ids = ["HG00097_hap1_hprc_r2_v1.0.1", "HG01891_mat_hprc_r2_v1.0.1", "HG01891_pat_hprc_r2_v1.0.1",
           "HG00097_hap2", "HG00097", "HG00097#1", "HG00097#1#", "  HG002_hap1  ", "", "#comment", "   ",
           "CHM13#0#chr2", "NA12878_hprc", "S00012_hap2_hprc_r2_v1.0.1", "sample_mat", "x_pat_hprc_junk"]
    return [[i, hfst.canonicalize_identifier(i)] for i in ids]


def window_cases(tmp):
    """Small synthetic windows: oracle similarity table -> reference scripts."""
    out = []
    specs = [(12, 2000, 11, None), (24, 5000, 12, None), (37, 3000, 13, None), (64, 4000, 14, None), (90, 6000, 15, 40)]
    for n, L, seed, ksites in specs:
        ws = synth.make_windows(n, L, 1, seed, n_sites_override=ksites)
        x = ws.dense(0)
        res = similarity.pairwise(x, ws.node_len[0])
        names = synth.haplotype_names(n, start=1000, end=1000 + L)
        path = os.path.join(tmp, f"w{n}.tsv")
        similarity.write_similarity_tsv(path, names, res)
        pops = ws.pops
        # pick the two largest populations present
        ids, cnts = np.unique(pops, return_counts=True)
        order = ids[np.argsort(-cnts)]
        ia = np.nonzero(pops == order[0])[0]
        ib = np.nonzero(pops == order[1])[0]
        pop_a = synth.assembly_names(ia)
        pop_b = synth.assembly_names(ib)
        iu = np.triu_indices(n, 1)
        idv = np.sort(res["identity"][iu])
        af_thr = [float(idv[len(idv) // 2]), float(idv[-max(1, len(idv) // 20)]), 1.0, float(idv[0])]
        case = table_cases(open(path).read(), pop_a, pop_b, tmp, f"wcase{n}", thresholds=(1.0,), expand=True,
                           lengths=(None, L), rounds=(None, 4), af_thresholds=af_thr)
        case.pop("tsv")   # reproducible from the stored matrix; keep the fixture small
        case.update({"n": n, "L": L, "m_pad": ws.m_pad, "names": names,
                     "x_bits": ws.x_bits[0].astype(np.uint32).tobytes().hex(),
                     "pitch_words": int(ws.x_bits.shape[2]),
                     "node_len": ws.node_len[0].astype(int).tolist(),
                     "pops": pops.astype(int).tolist(),
                     "idx_a": ia.astype(int).tolist(), "idx_b": ib.astype(int).tolist(),
                     "A": res["A"].astype(int).tolist(),
                     "I_upper": res["I"][iu].astype(int).tolist(),
                     "identity_upper": hx(res["identity"][iu].tolist()),
                     "S_all": similarity.segregating_nodes(x, ws.node_len[0])})
        out.append(case)
    return out


MESSY = """extra\tgroup.b\tgroup.a\testimated.identity\tjunk
1\tb#1#c:1-2\ta#1#c:1-2\t0.99\tx
2\tc#1#c:1-2\ta#1#c:1-2\t0.97\tx
3\tc#1#c:1-2\tb#1#c:1-2\t0.98\tx
4\td#2#c:1-2\ta#1#c:1-2\t0.5\tx
5\ta#1#c:1-2\tb#1#c:1-2\t0.995\tdup-last-wins
6\te#1#c:1-2\td#2#c:1-2\t1.0\tx
7\te#1#c:1-2\tc#1#c:1-2\t0.25\tx
"""


def main():
    assert refload.available(), "reference tree not found"
    with tempfile.TemporaryDirectory() as tmp:
        f6_tsv, pop_a, pop_b = extract_f6()
        gold = {
            "f6": table_cases(f6_tsv, pop_a, pop_b, tmp, "f6", thresholds=(1.0, 0.999, 0.99), lengths=(None, 50000, 100000)),
            "f6_cli": cli_cases(tmp, f6_tsv, pop_a, pop_b),
            "messy": table_cases(MESSY, ["a", "b", "c"], ["d#2", "e_hap1_hprc_zzz"], tmp, "messy",
                                 thresholds=(1.0,), expand=True, lengths=(None, 1000), rounds=(None, 1),
                                 af_thresholds=(0.98, 0.99, 1.0, 0.2)),
            "tajima": tajima_grid(),
            "canonical": canonical_cases(),
        }
        with open(os.path.join(HERE, "reference_outputs.json"), "w") as fh:
            json.dump(gold, fh, indent=1)
        with open(os.path.join(HERE, "windows.json"), "w") as fh:
            json.dump(window_cases(tmp), fh)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
