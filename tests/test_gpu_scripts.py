"""GPU parity of the drop-in script API (impop_b200.pica2 / hfst / tj_d / af) against the
golden vectors produced by the UNMODIFIED reference scripts (tests/golden/make_golden.py),
and of the auxiliary kernels against the CPU oracle.  Everything goes through the C ABI."""
import io
import os
import subprocess
import sys
from contextlib import redirect_stderr, redirect_stdout

import numpy as np
import pytest

from conftest import ROOT, rel_close, unhex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from impop_b200 import synth  # noqa: E402
from oracle import clib, popstats, similarity  # noqa: E402

TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    from impop_b200.runtime import default_context
    return default_context()


def _table(tmp_path, text, name="t.tsv"):
    p = tmp_path / name
    p.write_text(text if text.endswith("\n") else text + "\n")
    return str(p)


def _check_table_case(case, path):
    from impop_b200 import af, hfst, pica2
    sim, elements, pair_count = pica2.read_similarity_file(path)
    names = sorted(elements)
    for row in case["pica2"]:
        if not row["transitive"]:
            continue          # the reference's own answer depends on PYTHONHASHSEED there
        pi, pps = pica2.analyze_similarity_matrix(sim, elements, pair_count, threshold=row["threshold"],
                                                  sequence_length=row["L"], round_digits=row["round"])
        assert rel_close(pi, unhex(row["pi"]), TOL), row
        want_pps = unhex(row["pi_per_site"])
        if want_pps is None:
            assert pps is None
        else:
            assert rel_close(pps, want_pps, TOL), row
    sim2, all_sequences = hfst.read_similarity_file(path)
    if case.get("expand"):
        pa, miss_a = hfst.expand_population(case["pop_a"], all_sequences)
        pb, miss_b = hfst.expand_population(case["pop_b"], all_sequences)
        assert sorted(pa) == case["expanded"]["a"] and sorted(pb) == case["expanded"]["b"]
        assert sorted(miss_a) == case["expanded"]["missing_a"] and sorted(miss_b) == case["expanded"]["missing_b"]
    else:
        pa, pb = set(case["pop_a"]), set(case["pop_b"])
    for row in case["hfst"]:
        got = hfst.calculate_fst(sim2, pa, pb, sequence_length=row["L"], round_digits=row["round"])
        want = unhex(row["res"])
        for key in ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da"):
            assert rel_close(got[key], want[key], TOL), (row["L"], row["round"], key, got[key], want[key])
    w = unhex(case["diversity"]["within_a"])
    got = hfst.calculate_diversity(sim2, pa)
    assert rel_close(got[0], w[0], TOL) and got[1] == w[1] and got[2] == w[2]
    b = unhex(case["diversity"]["between"])
    got = hfst.calculate_diversity(sim2, pa, pb)
    assert rel_close(got[0], b[0], TOL) and got[1] == b[1] and got[2] == b[2]
    rows, samples = af.load_pairs(path)
    for row in case["af"]:
        summary = af.build_summary(af.cluster(rows, samples, row["threshold"]))
        want = row["summary"]
        assert len(summary) == len(want)
        for (cid, cnt, fr, mem), (wcid, wcnt, wfr, wmem) in zip(summary, want):
            assert cid == wcid and cnt == wcnt and mem == wmem and fr == unhex(wfr)
        fast = af.load_table_fast(path)                       # the command line's native reader path gives the same clusters
        if fast is not None:
            assert af.build_summary(af.cluster(None, None, row["threshold"], table=fast)) == summary
    return names


def test_f6_fixture(ctx, gold, tmp_path):
    """hudson/example_fst_methods.py:7-37, the reference's only fixture."""
    from impop_b200 import hfst, pica2
    case = gold["f6"]
    path = _table(tmp_path, case["tsv"])
    _check_table_case(case, path)
    sim, elements, pair_count = pica2.read_similarity_file(path)
    assert pair_count == 15 and len(elements) == 6
    res = hfst.calculate_fst(sim, set(case["pop_a"]), set(case["pop_b"]))
    assert rel_close(res["fst"], 0.9100000000000026, TOL) and rel_close(res["dxy"], 0.0050000000000000044, TOL)
    pi, pps = pica2.analyze_similarity_matrix(sim, elements, pair_count, threshold=1.0, sequence_length=100000)
    assert rel_close(pi, 0.0031799999999999975, TOL) and rel_close(pps, 3.1799999999999974e-08, TOL)
    # grouped: two cliques at 0.999 -> G1G2 term (1 - 0.995) * (3/6) * (3/6), pi = 6/5 * 0.0025
    pi, _ = pica2.analyze_similarity_matrix(sim, elements, pair_count, threshold=0.999)
    assert rel_close(pi, 6 / 5 * (2 * ((1 - 0.995) * 0.5 * 0.5)), TOL)


def test_messy_table(ctx, gold, tmp_path):
    """Reordered / extra columns, duplicate pair (last wins), absent pairs."""
    from impop_b200 import hfst
    case = gold["messy"]
    path = _table(tmp_path, case["tsv"])
    sim, _ = hfst.read_similarity_file(path)
    assert sim[("a#1#c:1-2", "b#1#c:1-2")] == 0.995
    _check_table_case(case, path)


def test_window_tables(ctx, gold_windows, tmp_path):
    """Synthetic windows: the device's own pairwise output -> TSV (repr precision) -> the drop-in API,
    against what the reference scripts returned for the oracle's TSV of the same window."""
    from impop_b200.engine import WindowBatch
    for case in gold_windows:
        n, pitch = case["n"], case["pitch_words"]
        bits = np.frombuffer(bytes.fromhex(case["x_bits"]), dtype=np.uint32).reshape(n, pitch)
        node_len = np.array(case["node_len"], dtype=np.uint32)
        batch = WindowBatch.from_windows(ctx, [(bits, node_len, np.full(n, 9, dtype=np.uint8), case["L"])])
        I, A, pi = batch.pairwise(0)
        ctx.check()
        I, A, pi = I.cpu().numpy(), A.cpu().numpy(), pi.cpu().numpy()
        U = A[:, None] + A[None, :] - I
        res = {"identity": 1.0 - pi, "I": I, "A": A, "J": I / np.maximum(U, 1)}
        path = str(tmp_path / f"w{n}.tsv")
        similarity.write_similarity_tsv(path, case["names"], res)
        _check_table_case(case, path)
        batch.close()


def test_tajima_grid(ctx, gold):
    from impop_b200 import tj_d
    rows = gold["tajima"]
    for row in rows:
        d, comps = tj_d.tajimas_d(row["n"], unhex(row["S"]), unhex(row["pi"]), return_components=True)
        want = unhex(row["D"])
        assert (d != d and want != want) or d == want, row       # bit-exact: same op order
        got = [comps.a1, comps.a2, comps.b1, comps.b2, comps.c1, comps.c2, comps.e1, comps.e2,
               comps.numerator, comps.denominator]
        for g, w in zip(got, unhex(row["parts"])):
            assert (g != g and w != w) or g == w
    ds = tj_d.tajimas_d_batch([r["n"] for r in rows], [unhex(r["S"]) for r in rows], [unhex(r["pi"]) for r in rows])
    for d, row in zip(ds, rows):
        want = unhex(row["D"])
        assert (d != d and want != want) or d == want
    with pytest.raises(ValueError):
        tj_d.tajimas_d(1, 1.0, 0.1)
    with pytest.raises(ValueError):
        tj_d.tajimas_d(10, -1.0, 0.1)


def test_canonical_prefix(gold):
    from impop_b200 import hfst
    for ident, want in gold["canonical"]:
        assert hfst.canonicalize_identifier(ident) == want, ident


def _run(script, *argv):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *argv], capture_output=True, text=True,
                          timeout=300)
    return proc.stdout, proc.stderr, proc.returncode


def test_cli_known_answers(ctx, gold, tmp_path):
    """Appendix A: stdout strings and exit codes of the reference CLIs on the F6 fixture."""
    case, cli = gold["f6"], gold["f6_cli"]
    path = _table(tmp_path, case["tsv"], "example_similarities.tsv")
    logs = str(tmp_path / "logs")
    out, _, code = _run("pica2.py", path, "-t", "0.999", "-d", logs)
    assert code == 0 and out == cli["pica2_t0999"]["stdout"]
    assert os.path.exists(os.path.join(logs, "example_similarities.log"))
    out, _, code = _run("pica2.py", path, "-t", "0.999", "-l", "1000000", "-r", "5", "-d", logs)
    assert code == 0 and out == cli["pica2_t0999_l_r5"]["stdout"]
    pa, pb = tmp_path / "pop_A.txt", tmp_path / "pop_B.txt"
    pa.write_text("\n".join(case["pop_a"]) + "\n")
    pb.write_text("\n".join(case["pop_b"]) + "\n")
    out, err, code = _run("h-fst.py", path, "-a", str(pa), "-b", str(pb), "-d", logs)
    assert code == cli["hfst_f6"]["code"] == 1 and "No valid sequences" in err      # ids lack '#': a-3
    out, _, code = _run("tj_d.py", "-n", "446", "-p", "0.59146123", "-S", "20", "--show-components")
    assert code == 0 and out == cli["tjd_doc"]["stdout"]
    out, _, code = _run("tj_d.py", "-n", "446", "-p", "0.5", "-S", "0")
    assert out == cli["tjd_s0"]["stdout"]
    out, _, code = _run("af.py", "--input", path, "--threshold", "0.9995")
    assert code == 0 and out.split() == cli["af_09995"]["stdout"].split()
    out, _, code = _run("pica2.py", str(tmp_path / "missing.tsv"))
    assert code == 1 and "File not found" in out


def test_hfst_cli_on_pansn_window(ctx, gold_windows, tmp_path):
    """h-fst.py CLI with assembly-name population lists (exercises canonicalize/expand, a-3) vs golden."""
    case = gold_windows[1]
    n, pitch = case["n"], case["pitch_words"]
    bits = np.frombuffer(bytes.fromhex(case["x_bits"]), dtype=np.uint32).reshape(n, pitch)
    x = similarity.unpack_bits(bits, case["m_pad"])
    res = similarity.pairwise(x, np.array(case["node_len"], dtype=np.uint32))
    path = str(tmp_path / "win.tsv")
    similarity.write_similarity_tsv(path, case["names"], res)
    pa, pb = tmp_path / "a.txt", tmp_path / "b.txt"
    pa.write_text("\n".join(case["pop_a"]) + "\n")
    pb.write_text("# comment\n\n" + "\n".join(case["pop_b"]) + "\n")
    out, err, code = _run("h-fst.py", path, "-a", str(pa), "-b", str(pb), "-l", str(case["L"]), "-d", str(tmp_path))
    assert code == 0, err
    want = unhex([r for r in case["hfst"] if r["L"] == case["L"] and r["round"] is None][0]["res"])
    got = [float(v) for v in out.split("\t")]
    for g, key in zip(got, ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da")):
        assert f"{want[key]:.8f}" == f"{g:.8f}"
    assert os.path.exists(tmp_path / "win_fst.log")


# ----------------------------------------------------------------------------------------------
# auxiliary kernels vs the oracle
# ----------------------------------------------------------------------------------------------
def test_pack_bits(ctx):
    rng = np.random.default_rng(3)
    for n, m in ((1, 1), (5, 31), (7, 32), (9, 33), (130, 1000), (466, 1009)):
        dense = (rng.random((n, m)) < 0.4).astype(np.uint8) * rng.integers(1, 255, size=(n, m), dtype=np.uint8)
        got = ctx.pack_bits(torch.from_numpy(dense).to(ctx.torch_device))
        ctx.check()
        want = similarity.pack_bits((dense != 0).astype(np.uint8))
        assert (got.cpu().numpy().view(np.uint32) == want).all(), (n, m)


def test_site_counts(ctx):
    """BASELINE config 4 shape (466 haplotypes, 5 + 1 panels) and odd shapes; counts and freq bit-exact."""
    from impop_b200 import af
    sites, masks = synth.make_site_matrix(50000, 466, seed=0xB200 + 4)
    for mk in (masks[:5], masks, masks[:1], masks[:3]):
        counts, freq = af.site_allele_counts(sites, mk, ctx=ctx)
        want_c, want_f = clib.site_counts(sites, mk)
        assert (counts == want_c).all() and (freq == want_f).all()
    for n, sites_n in ((70, 1000), (1000, 777), (64, 1)):
        s2, m2 = synth.make_site_matrix(sites_n, n, seed=n)
        counts, freq = af.site_allele_counts(s2, m2, ctx=ctx)
        want_c, want_f = clib.site_counts(s2, m2)
        assert (counts == want_c).all() and (freq == want_f).all()
    counts, freq = af.site_allele_counts(sites[:0], masks[:5], ctx=ctx)
    assert counts.shape == (0, 5)


def test_reduce_identity_with_missing_pairs(ctx):
    """TSV mode on a random table with absent pairs and overlapping label classes vs the Python oracle."""
    rng = np.random.default_rng(11)
    n = 300
    names = [f"h{i:04d}" for i in range(n)]
    mat = 1.0 - rng.random((n, n)) * 0.01
    mat = np.triu(mat, 1) + np.triu(mat, 1).T
    gone = np.triu(rng.random((n, n)) < 0.05, 1)
    mat[gone | gone.T] = np.nan
    np.fill_diagonal(mat, np.nan)
    ia, ib = list(range(0, 120)), list(range(100, 250))          # overlapping on purpose at the label level
    lab = np.zeros(n, dtype=np.uint8)
    lab[:] |= 1
    lab[ia] |= 2
    lab[ib] |= 4
    stats, counts, _ = ctx.reduce_identity(torch.from_numpy(mat).to(ctx.torch_device),
                                           torch.from_numpy(lab).to(ctx.torch_device), None, length=5000, seg_sites=37.0)
    ctx.check()
    stats, counts = stats.cpu().numpy(), counts.cpu().numpy()
    pa, ca, _ = popstats.mean_diversity(mat, ia)
    pb, cb, _ = popstats.mean_diversity(mat, ib)
    assert counts[4] == ca and counts[5] == cb
    assert rel_close(stats[2] * 5000, pa, 1e-11) and rel_close(stats[3] * 5000, pb, 1e-11)
    pi, pps = popstats.pica2_pi(mat, names, 1.0, 5000)
    # with absent pairs pica2's n/(n-1) * sum uses the present pairs only -- same on the device
    assert rel_close(stats[0], pi, TOL) and rel_close(stats[1], pps, TOL)
    d, _ = popstats.tajimas_d(n, 37.0, pps)
    assert stats[9] == d or rel_close(stats[9], d, TOL)         # plain 1e-12 relative (north star)


def test_greedy_groups_transitive_and_threshold_edges(ctx):
    """Cliques: identical to the reference whatever its hash seed; strict '>' at the threshold."""
    from impop_b200 import pica2
    from impop_b200.tables import SimilarityTable
    rng = np.random.default_rng(5)
    n, k = 60, 7
    member = rng.integers(0, k, size=n)
    names = [f"s{i:03d}" for i in range(n)]
    between = 0.99 - rng.random((k, k)) * 0.01
    between = np.minimum(between, between.T)
    mat = between[member][:, member]
    mat[member[:, None] == member[None, :]] = 0.9995
    np.fill_diagonal(mat, np.nan)
    tab = SimilarityTable(names, mat)
    for thr in (0.999, 0.9995, 1.0, 0.5):
        got = pica2.analyze_similarity_matrix(tab, set(names), n * (n - 1) // 2, threshold=thr, sequence_length=1000, ctx=ctx)
        want = popstats.pica2_pi(mat, names, thr, 1000)
        assert rel_close(got[0], want[0], TOL) and rel_close(got[1], want[1], TOL), thr
    group, weight = ctx.greedy_groups(tab.device(ctx), 0.999)
    ctx.check()
    g = group.cpu().numpy()
    for a in range(n):
        for b in range(n):
            assert (g[a] == g[b]) == (member[a] == member[b])
    assert abs(float(weight.cpu().numpy().sum()) - 1.0) < 1e-12


def test_cluster_random_graph(ctx):
    from impop_b200 import af
    rng = np.random.default_rng(8)
    n = 200
    names = [f"x{i:03d}#1#c" for i in range(n)]
    rows = []
    for i in range(n):
        for j in range(i + 1, n):
            v = 1.0 if rng.random() < 0.004 else float(0.9 + 0.09 * rng.random())
            rows.append((names[i] + ":1-2", names[j] + ":1-2", v))
    rows_af = [(a.split(":")[0], b.split(":")[0], v) for a, b, v in rows]
    mat = np.full((n, n), np.nan)
    for a, b, v in rows_af:
        i, j = names.index(a), names.index(b)
        mat[i, j] = mat[j, i] = v
    for thr in (1.0, 0.985, 0.95):
        got = af.cluster(rows_af, names, thr, ctx=ctx)
        want = popstats.af_clusters(mat, names, thr)
        assert got == want, thr
