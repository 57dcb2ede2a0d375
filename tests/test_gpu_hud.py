"""Row f-3 on the GPU: the hud.py drop-in (grouped and direct methods, function level and CLI) against outputs of the
unmodified reference stored in tests/golden/hud_grouped.json."""
import json
import os

import pytest

from conftest import unhex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hud_grouped.json")))
TOL = 1e-12


def _close(got, want, tag):
    scale = max(abs(want["dxy"]), abs(want["pi_xy"]))
    for k in ("pi_a", "pi_b", "pi_xy", "dxy"):
        assert abs(got[k] - want[k]) <= TOL * max(abs(want[k]), 1e-300), (tag, k, got[k], want[k])
    assert abs(got["da"] - want["da"]) <= 4 * TOL * scale, (tag, "da", got["da"], want["da"])
    if want["dxy"] > 0:
        assert abs(got["fst"] - want["fst"]) <= 8 * TOL * scale / abs(want["dxy"]), (tag, "fst", got["fst"], want["fst"])
    else:
        assert got["fst"] == 0.0


@pytest.mark.parametrize("tag", sorted(GOLD))
def test_hud_functions_match_the_reference(tag, tmp_path):
    from impop_b200 import hud
    g = GOLD[tag]
    path = tmp_path / "t.tsv"
    path.write_text(g["tsv"])
    sims, seqs = hud.read_similarity_file(str(path))
    for case in g["cases"]:
        got = hud.calculate_fst(sims, set(g["pop_a"]), set(g["pop_b"]), sequence_length=case["L"], round_digits=case["round"],
                                method="grouped", threshold=case["threshold"])
        _close(got, {k: unhex(v) for k, v in case["grouped"].items()}, (tag, "grouped", case["threshold"], case["round"], case["L"]))
        got = hud.calculate_fst(sims, set(g["pop_a"]), set(g["pop_b"]), sequence_length=case["L"], round_digits=case["round"],
                                method="direct")
        _close(got, {k: unhex(v) for k, v in case["direct"].items()}, (tag, "direct", case["round"], case["L"]))


def test_hud_cli_and_groups(tmp_path, capsys):
    from impop_b200 import hud
    g = GOLD["blocky0"]
    (tmp_path / "t.tsv").write_text(g["tsv"])
    (tmp_path / "a.txt").write_text("\n".join(g["pop_a"]) + "\n")
    (tmp_path / "b.txt").write_text("\n".join(g["pop_b"]) + "\n")
    case = next(c for c in g["cases"] if c["threshold"] == 0.9995 and c["round"] is None and c["L"] == 50000)
    rc = hud.main([str(tmp_path / "t.tsv"), "-a", str(tmp_path / "a.txt"), "-b", str(tmp_path / "b.txt"), "-l", "50000",
                   "-m", "grouped", "-t", "0.9995", "-d", str(tmp_path / "logs")])
    assert rc == 0
    out = capsys.readouterr().out.strip().split("\t")
    want = {k: unhex(v) for k, v in case["grouped"].items()}
    assert out == [f"{want[k]:.8f}" for k in ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da")]
    log = (tmp_path / "logs" / "t_fst.log").read_text()
    assert "Method: grouped" in log and "Grouping threshold: 0.9995" in log
    sims, seqs = hud.read_similarity_file(str(tmp_path / "t.tsv"))
    groups = hud.group_sequences(sims, set(g["pop_a"]), threshold=0.9995)
    assert groups == sorted(groups) and sum(len(x) for x in groups) == len(g["pop_a"])
    assert all(x == sorted(x) for x in groups) and len(groups) == 3        # clusters 0, 1 and the rest of cluster 2
