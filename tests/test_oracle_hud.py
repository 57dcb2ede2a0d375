"""Row f-3: the oracle's deterministic restatement of hudson/hud.py's grouped method against outputs of the unmodified
reference (tests/golden/hud_grouped.json, made by tests/golden/make_golden_hud.py on tables where the reference's
hash-order-dependent grouping has a single outcome)."""
import json
import os

import numpy as np
import pytest

from conftest import unhex
from oracle import popstats

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hud_grouped.json")))


@pytest.mark.parametrize("tag", sorted(GOLD))
def test_grouped_and_direct_match_the_reference(tag, tmp_path):
    g = GOLD[tag]
    path = tmp_path / "t.tsv"
    path.write_text(g["tsv"])
    names, mat, _ = popstats.parse_similarity_tsv(str(path))
    for case in g["cases"]:
        got = popstats.hud_fst_grouped(mat, names, set(g["pop_a"]), set(g["pop_b"]), sequence_length=case["L"],
                                       round_digits=case["round"], threshold=case["threshold"])
        want = {k: unhex(v) for k, v in case["grouped"].items()}
        for k in ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da"):
            assert got[k] == want[k] or (np.isnan(got[k]) and np.isnan(want[k])), (tag, case["threshold"], case["round"], case["L"], k, got[k], want[k])
        direct = popstats.hudson_fst(mat, names, set(g["pop_a"]), set(g["pop_b"]), sequence_length=case["L"], round_digits=case["round"])
        wd = {k: unhex(v) for k, v in case["direct"].items()}
        for k in ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da"):
            assert abs(direct[k] - wd[k]) <= 1e-12 * max(abs(wd[k]), 1e-300) or direct[k] == wd[k], (tag, "direct", k)
