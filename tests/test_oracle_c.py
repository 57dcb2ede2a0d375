"""Pin the plain-C oracle (oracle/csrc/oracle_impop.c) to the numpy/Python oracle and the goldens."""
import numpy as np
import pytest

from conftest import rel_close, unhex
from impop_b200 import synth
from oracle import clib, popstats, similarity

LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = 1, 2, 4, 8


def _labels(n, ia, ib, subset=None, seg=None):
    lab = np.zeros(n, dtype=np.uint8)
    lab[list(range(n)) if subset is None else list(subset)] |= LAB_SUBSET
    lab[list(ia)] |= LAB_A
    lab[list(ib)] |= LAB_B
    lab[list(range(n)) if seg is None else list(seg)] |= LAB_SEG
    return lab


def test_c_pairwise_matches_numpy_and_golden(gold_windows):
    for case in gold_windows:
        n, pitch = case["n"], case["pitch_words"]
        bits = np.frombuffer(bytes.fromhex(case["x_bits"]), dtype=np.uint32).reshape(n, pitch)
        node_len = np.array(case["node_len"], dtype=np.uint32)
        iu = np.triu_indices(n, 1)
        for use_lut in (False, True):
            A, I, pi = clib.window_pairwise(bits, case["m_pad"], node_len, use_lut=use_lut)
            assert A.tolist() == case["A"]
            assert I[iu].tolist() == case["I_upper"]
            ident = np.array(unhex(case["identity_upper"]))
            assert (pi[iu] == 1.0 - ident).all()          # bit-exact: same op order


def test_c_window_stats_vs_reference_golden(gold_windows):
    """C oracle fused stats vs the reference's own pica2 / h-fst outputs stored in the goldens."""
    for case in gold_windows:
        n, pitch, L = case["n"], case["pitch_words"], case["L"]
        bits = np.frombuffer(bytes.fromhex(case["x_bits"]), dtype=np.uint32).reshape(n, pitch)
        node_len = np.array(case["node_len"], dtype=np.uint32)
        for use_len in (None, L):
            stats, counts = clib.window_stats(bits, case["m_pad"], node_len, _labels(n, case["idx_a"], case["idx_b"]), use_len or 0)
            pica = [r for r in case["pica2"] if r["threshold"] == 1.0 and r["L"] == use_len and r["round"] is None][0]
            assert rel_close(stats[0], unhex(pica["pi"]))
            if use_len:
                assert rel_close(stats[1], unhex(pica["pi_per_site"]))
            else:
                assert np.isnan(stats[1])
            hf = unhex([r for r in case["hfst"] if r["L"] == use_len and r["round"] is None][0]["res"])
            for col, key in ((2, "pi_a"), (3, "pi_b"), (4, "pi_xy"), (5, "dxy"), (6, "da"), (7, "fst")):
                assert rel_close(stats[col], hf[key]), (key, stats[col], hf[key])
            assert counts[7] == case["S_all"] and stats[8] == case["S_all"]
            assert counts[0] == n and counts[3] == n * (n - 1) // 2
            # Tajima's D of the window equals the restated formula on (n, S, pi used)
            pi_used = stats[1] if use_len else stats[0]
            d, parts = popstats.tajimas_d(n, float(case["S_all"]), pi_used)
            assert (np.isnan(d) and np.isnan(stats[9])) or d == stats[9]
            assert stats[10] == parts.a1 and stats[11] == parts.e1 and stats[12] == parts.e2


def test_c_tajima_bit_exact(gold):
    for row in gold["tajima"]:
        d, parts = clib.tajimas_d(row["n"], unhex(row["S"]), unhex(row["pi"]))
        want = unhex(row["D"])
        assert (d != d and want != want) or d == want
        for g, w in zip(parts.tolist(), unhex(row["parts"])):
            assert (g != g and w != w) or g == w


@pytest.mark.parametrize("n,m,seed", [(2, 1, 1), (3, 33, 2), (17, 64, 3), (40, 300, 4), (129, 257, 5)])
def test_c_random_windows(n, m, seed):
    """Random matrices incl. heavy nodes (len > 255, > 65535), zero-length nodes and ragged m."""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, m)) < 0.6).astype(np.uint8)
    node_len = rng.integers(0, 60, size=m).astype(np.uint32)
    heavy = rng.random(m) < 0.1
    node_len[heavy] = rng.integers(256, 200000, size=int(heavy.sum()))
    if n > 2:
        x[1] = x[0]                     # identical haplotypes -> pi == 0
        x[2] = 0                        # empty path
    bits = similarity.pack_bits(x)
    ref = similarity.pairwise(x, node_len)
    A, I, pi = clib.window_pairwise(bits, m, node_len)
    assert (A == ref["A"]).all() and (I == ref["I"]).all()
    off = ~np.eye(n, dtype=bool)
    assert (pi[off] == ref["pi"][off]).all()
    ia = list(range(0, n, 3))
    ib = list(range(1, n, 3))
    sub = list(range(0, n, 2)) if n > 3 else list(range(n))
    lab = _labels(n, ia, ib, subset=sub, seg=sub)
    stats, counts = clib.window_stats(bits, m, node_len, lab, 1234)
    names = [f"h{i:04d}" for i in range(n)]
    p = ref["pi"]
    if len(sub) >= 2:
        want = popstats.pica2_pi(1.0 - p[np.ix_(sub, sub)], [names[i] for i in sub], 1.0, 1234)
        assert rel_close(stats[0], want[0]) and rel_close(stats[1], want[1])
    hf = popstats.hudson_fst(1.0 - p, names, {names[i] for i in ia}, {names[i] for i in ib}, 1234)
    for col, key in ((2, "pi_a"), (3, "pi_b"), (4, "pi_xy"), (5, "dxy"), (6, "da"), (7, "fst")):
        assert rel_close(stats[col], hf[key]), key
    assert counts[7] == similarity.segregating_nodes(x, node_len, rows=sub)


def test_c_site_counts():
    sites, masks = synth.make_site_matrix(2000, 466, seed=7)
    counts, freq = clib.site_counts(sites, masks)
    want_c, want_f = popstats.site_allele_counts(sites, masks)
    assert (counts == want_c).all() and (freq == want_f).all()
