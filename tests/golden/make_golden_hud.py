#!/usr/bin/env python3
"""Regenerate tests/golden/hud_grouped.json by running the UNMODIFIED reference hudson/hud.py (grouped and direct
methods) -- build container only (needs /root/reference).  Only tables on which `sim > threshold` is an equivalence
relation inside each population are kept: there hud.group_sequences (hud.py:64-84, seeded by set.pop()) has one
possible outcome, so the stored numbers do not depend on PYTHONHASHSEED."""
from __future__ import annotations

import io
import json
import os
import sys
from contextlib import redirect_stderr

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import extract_f6, hx  # noqa: E402
from oracle import popstats, refload  # noqa: E402


def transitive_within(mat, idx, thr):
    sub = np.nan_to_num(mat[np.ix_(idx, idx)], nan=-np.inf) > thr
    np.fill_diagonal(sub, True)
    reach = sub.copy()
    for _ in range(len(idx)):
        reach = reach | ((reach.astype(np.int64) @ reach.astype(np.int64)) > 0)
    return bool((reach == sub).all())


def blocky_table(seed, sizes, within=(0.99951, 0.99999), between=(0.95, 0.9985)):
    """Clusters of near-identical sequences: identity inside a cluster above any grouping threshold <= 0.9995."""
    rng = np.random.default_rng(seed)
    n = sum(sizes)
    cl = np.repeat(np.arange(len(sizes)), sizes)
    names = [f"S{i:03d}#{1 + i % 2}#ctg{i}:1000-51000" for i in range(n)]
    rows = ["group.a\tgroup.b\testimated.identity"]
    for i in range(n):
        for j in range(i + 1, n):
            lo, hi = within if cl[i] == cl[j] else between
            rows.append(f"{names[i]}\t{names[j]}\t{rng.uniform(lo, hi)!r}")
    return "\n".join(rows) + "\n", names, cl


def cases_for(hud, tsv_text, pop_a, pop_b, tmp, tag, thresholds, lengths=(None, 50000), rounds=(None, 4)):
    path = os.path.join(tmp, f"{tag}.tsv")
    with open(path, "w") as fh:
        fh.write(tsv_text)
    names, mat, _ = popstats.parse_similarity_tsv(path)
    where = {s: i for i, s in enumerate(names)}
    out = {"tsv": tsv_text, "pop_a": sorted(pop_a), "pop_b": sorted(pop_b), "cases": []}
    with redirect_stderr(io.StringIO()):
        sims, seqs = hud.read_similarity_file(path)
        for thr in thresholds:
            for r in rounds:
                rm = popstats.py_round_matrix(mat, r)
                ok = transitive_within(rm, [where[s] for s in sorted(pop_a)], thr) and \
                    transitive_within(rm, [where[s] for s in sorted(pop_b)], thr)
                if not ok:
                    continue
                for L in lengths:
                    res = hud.calculate_fst(sims, set(pop_a), set(pop_b), sequence_length=L, round_digits=r,
                                            method="grouped", threshold=thr)
                    direct = hud.calculate_fst(sims, set(pop_a), set(pop_b), sequence_length=L, round_digits=r, method="direct")
                    out["cases"].append({"threshold": thr, "round": r, "L": L, "grouped": hx(res), "direct": hx(direct)})
    return out


def main():
    import tempfile
    assert refload.available(), "reference tree not found"
    hud = refload.load("hud")
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        f6_tsv, pa, pb = extract_f6()
        gold["f6"] = cases_for(hud, f6_tsv if f6_tsv.endswith("\n") else f6_tsv + "\n", pa, pb, tmp, "f6",
                               thresholds=(0.999, 0.996, 0.9995, 1.0))
        for k, (seed, sizes, na) in enumerate([(1, (4, 3, 5, 2, 6), 3), (2, (1, 1, 7, 3, 2, 2, 4), 4), (3, (10, 8), 1)]):
            tsv, names, cl = blocky_table(seed, sizes)
            pop_a = [s for s, c in zip(names, cl) if c < na]
            pop_b = [s for s, c in zip(names, cl) if c >= na]
            # move part of one cluster across so that a cluster is split between the populations
            pop_b.append(pop_a.pop())
            gold[f"blocky{k}"] = cases_for(hud, tsv, pop_a, pop_b, tmp, f"blocky{k}", thresholds=(0.9995, 0.999, 0.9, 1.0))
    assert all(v["cases"] for v in gold.values())
    with open(os.path.join(HERE, "hud_grouped.json"), "w") as fh:
        json.dump(gold, fh)
    print({k: len(v["cases"]) for k, v in gold.items()})


if __name__ == "__main__":
    main()
