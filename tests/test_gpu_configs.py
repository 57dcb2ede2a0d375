"""The other BASELINE.json configurations as parity cases (the bench line is configs[1]):
config 1 (90 haplotypes, 100 kb, TSV mode and matrix mode), config 3 (20 kb windows, Tajima's D),
config 4 (per-site allele frequencies at scale), config 5 (10 000 haplotypes, 200 kb, tile grid split)."""
import numpy as np
import pytest

from conftest import rel_close

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from impop_b200 import synth  # noqa: E402
from oracle import clib, popstats, similarity  # noqa: E402
from oracle.compare import row_mismatches  # noqa: E402

TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    from impop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _labels(n, ia, ib, subset=None, seg=None):
    lab = np.zeros(n, dtype=np.uint8)
    lab[list(range(n)) if subset is None else list(subset)] |= 1
    lab[list(ia)] |= 2
    lab[list(ib)] |= 4
    lab[list(range(n)) if seg is None else list(seg)] |= 8
    return lab


def test_config1_edar_window_tsv_and_matrix_mode(ctx, tmp_path):
    """pica2 on one EDAR-shaped 100 kb window of 90 haplotypes: TSV mode (drop-in API on the table the
    oracle writes) and matrix mode (fused kernel) give the same pi as the oracle, threshold 1.0."""
    from impop_b200 import pica2
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(90, 100000, 1, seed=0xB200 + 0)
    x = ws.dense(0)
    res = similarity.pairwise(x, ws.node_len[0])
    names = synth.haplotype_names(90, "chr2", 108894471, 108994471)
    path = str(tmp_path / "edar.tsv")
    similarity.write_similarity_tsv(path, names, res)
    sim, elements, rows = pica2.read_similarity_file(path)
    assert rows == 4005
    pi_tsv, pps_tsv = pica2.analyze_similarity_matrix(sim, elements, rows, threshold=1.0, sequence_length=100000)
    want = popstats.pica2_pi(1.0 - res["pi"], sorted(names), 1.0, 100000)
    # names sort like indices here, so the dense matrices agree element-wise
    assert rel_close(pi_tsv, want[0], TOL) and rel_close(pps_tsv, want[1], TOL)
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, _labels(90, [], []), ws.length)
    stats, _ = batch.stats(0)
    ctx.check()
    st = stats.cpu().numpy()[0]
    assert rel_close(st[0], want[0], TOL) and rel_close(st[1], want[1], TOL)
    batch.close()


def test_config3_tajima_20kb_windows(ctx):
    """tj_d genome-wide shape: 20 kb windows (m = 403 -> 512), 466 haplotypes; S, pi, D per window."""
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(466, 20000, 24, seed=0xB200 + 3)
    lab = _labels(466, [], [])
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    stats, counts = batch.stats(0)
    ctx.check()
    st, ct = stats.cpu().numpy(), counts.cpu().numpy()
    want_s, want_c = clib.batch_stats(batch.n, batch.m, batch.pitch_words, batch.x_off, batch.len_off, batch.lab_off,
                                      batch.length, ws.x_bits, ws.node_len, lab, 4)
    assert (ct == want_c).all()
    for w in range(ws.windows):
        assert not row_mismatches(st[w], want_s[w], TOL), w
        d, parts = popstats.tajimas_d(466, float(want_c[w][7]), float(want_s[w][1]))   # run_tajd.sh: per-site pi, absolute S
        assert st[w][9] == d or rel_close(st[w][9], d, TOL)      # plain 1e-12 relative (north star)
        assert st[w][10] == parts.a1 and st[w][11] == parts.e1 and st[w][12] == parts.e2
    batch.close()


def test_config4_site_frequencies_at_scale(ctx):
    """2 x 10^6 sites x 466 haplotypes x 5 panels: counts sum to the popcount of each row restricted to the
    panels, frequencies are count / |panel| bit-exactly, and a random sample of sites matches the oracle."""
    sites, masks = synth.make_site_matrix(2_000_000, 466, seed=0xB200 + 4)
    dev = ctx.torch_device
    ds = torch.from_numpy(sites.view(np.int64)).to(dev)
    dm = torch.from_numpy(masks[:5].view(np.int64)).to(dev)
    counts, freq = ctx.site_counts(ds, dm)
    ctx.check()
    counts, freq = counts.cpu().numpy(), freq.cpu().numpy()
    idx = np.random.default_rng(0).choice(sites.shape[0], 20000, replace=False)
    want_c, want_f = clib.site_counts(sites[idx], masks[:5])
    assert (counts[idx] == want_c).all() and (freq[idx] == want_f).all()
    union = np.bitwise_or.reduce(masks[:5], axis=0)
    lut = np.array([bin(v).count("1") for v in range(256)], dtype=np.int64)
    tot = lut[(sites & union[None, :]).view(np.uint8)].reshape(sites.shape[0], -1).sum(axis=1)
    assert (counts.sum(axis=1) == tot).all()                     # panels are disjoint
    sizes = np.array([140, 88, 100, 60, 72], dtype=np.float64)
    assert (freq == counts / sizes[None, :]).all()


def test_config5_scale_up_10000_haplotypes(ctx):
    """One 200 kb window of 10 000 haplotypes (m = 5 875 -> 5 888, 3 160 tiles in the upper triangle):
    (1) pi over a 600-haplotype SUBSET of the big window equals the oracle on the extracted rows,
    (2) the tile grid dealt to 2 / 4 ranks and re-assembled equals the single-GPU run bit for bit in the
        counts and to 1e-12 in the statistics, (3) two-panel Fst columns satisfy pi_xy = (pi_a + pi_b) / 2."""
    from impop_b200.engine import WindowBatch
    n = 10000
    ws = synth.make_windows(n, 200000, 1, seed=0xB200 + 5, chunk=1)
    rng = np.random.default_rng(5)
    sub = np.sort(rng.choice(n, 600, replace=False))
    pa, pb = sub[:300], sub[300:]
    lab = np.zeros(n, dtype=np.uint8)
    lab[sub] |= 1 | 8
    lab[pa] |= 2
    lab[pb] |= 4
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    stats, counts = batch.stats(0)
    ctx.check()
    st, ct = stats.cpu().numpy()[0], counts.cpu().numpy()[0]
    lab_sub = np.full(600, 9, dtype=np.uint8)
    lab_sub[:300] |= 2
    lab_sub[300:] |= 4
    want_s, want_c = clib.window_stats(np.ascontiguousarray(ws.x_bits[0][sub]), ws.m_pad, ws.node_len[0], lab_sub, ws.length)
    assert (ct == want_c).all()
    assert not row_mismatches(st, want_s, TOL)
    # all 10 000 haplotypes in two panels of 5 000: split grid == single run
    lab2 = np.full(n, 9, dtype=np.uint8)
    lab2[:5000] |= 2
    lab2[5000:] |= 4
    batch.labels.copy_(torch.from_numpy(lab2))
    s1, c1 = batch.stats(0)
    ctx.check()
    assert batch.items == sum((n - 128 * bi + 255) // 256 for bi in range((n + 127) // 128))
    for world in (2, 4):
        parts = torch.stack([batch.window_sums(r, world) for r in range(world)]).contiguous()
        s2, c2 = batch.finalize(parts)
        ctx.check()
        assert torch.equal(c1, c2)
        assert not row_mismatches(s2.cpu().numpy()[0], s1.cpu().numpy()[0], TOL)
    s = s1.cpu().numpy()[0]
    assert c1.cpu().numpy()[0].tolist()[:7] == [n, 5000, 5000, n * (n - 1) // 2, 5000 * 4999 // 2, 5000 * 4999 // 2, 25_000_000]
    assert rel_close(s[4], 0.5 * (s[2] + s[3]), 1e-15)
    assert 0.0 < s[7] < 1.0 and s[5] > s[4] > 0.0                   # structured panels: Dxy > pi_xy, 0 < Fst < 1
    # a random handful of pairs of the big window, straight from the materialising call
    batch.close()
