"""GPU parity: the CUDA window path (through the C ABI) against the CPU oracle.

Integers (I, A, S, counts) bit-exact; pi_ij bit-exact (same fp64 op order); summed
statistics within 1e-12 relative (north star tolerance) -- they differ only by summation order.
"""
import numpy as np
import pytest

from conftest import rel_close, unhex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from impop_b200 import synth  # noqa: E402
from oracle import clib, similarity  # noqa: E402
from oracle.compare import row_mismatches  # noqa: E402

LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = 1, 2, 4, 8
TOL = 1e-12
ALGOS = [0, 1]  # tcgen05, simt


@pytest.fixture(scope="module")
def ctx():
    from impop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _labels(n, ia, ib, subset=None, seg=None):
    lab = np.zeros(n, dtype=np.uint8)
    lab[list(range(n)) if subset is None else list(subset)] |= LAB_SUBSET
    lab[list(ia)] |= LAB_A
    lab[list(ib)] |= LAB_B
    lab[list(range(n)) if seg is None else list(seg)] |= LAB_SEG
    return lab


def _random_window(n, m, seed, heavy_frac=0.1, max_len=200000, density=0.6):
    rng = np.random.default_rng(seed)
    x = (rng.random((n, m)) < density).astype(np.uint8)
    node_len = rng.integers(0, 60, size=m).astype(np.uint32)
    heavy = rng.random(m) < heavy_frac
    node_len[heavy] = rng.integers(255, max_len, size=int(heavy.sum()))
    if n > 2:
        x[1] = x[0]
        x[2] = 0
    if n > 4:
        x[4] = 1 - x[3]                # empty intersection
    return x, node_len


def _assert_stats(got, want, ctxmsg=""):
    """1e-12 relative per column; da / fst / D relative to their operands' scale (oracle/compare.py)."""
    bad = row_mismatches(got, want, TOL)
    assert not bad, (ctxmsg, bad)


@pytest.mark.parametrize("algo", ALGOS)
def test_pairwise_golden_windows(ctx, gold_windows, algo):
    """Committed fixtures: I, A bit-exact and pi == 1 - identity bit-exact."""
    from impop_b200.engine import WindowBatch
    wins = []
    for case in gold_windows:
        n, pitch = case["n"], case["pitch_words"]
        bits = np.frombuffer(bytes.fromhex(case["x_bits"]), dtype=np.uint32).reshape(n, pitch)
        wins.append((bits, np.array(case["node_len"], dtype=np.uint32), _labels(n, case["idx_a"], case["idx_b"]), case["L"]))
    batch = WindowBatch.from_windows(ctx, wins)
    for w, case in enumerate(gold_windows):
        n = case["n"]
        I, A, pi = batch.pairwise(w, algo)
        ctx.check()
        iu = np.triu_indices(n, 1)
        assert A.cpu().tolist() == case["A"]
        assert I.cpu().numpy()[iu].tolist() == case["I_upper"]
        ident = np.array(unhex(case["identity_upper"]))
        assert (pi.cpu().numpy()[iu] == 1.0 - ident).all()
    stats, counts = batch.stats(algo)
    ctx.check()
    stats, counts = stats.cpu().numpy(), counts.cpu().numpy()
    for w, (case, win) in enumerate(zip(gold_windows, wins)):
        want_s, want_c = clib.window_stats(win[0], case["m_pad"], win[1], win[2], case["L"])
        assert (counts[w] == want_c).all()
        _assert_stats(stats[w], want_s, f"golden window {w}")
        # and directly against the reference's own pica2 / h-fst outputs stored in the fixture
        pica = [r for r in case["pica2"] if r["threshold"] == 1.0 and r["L"] == case["L"] and r["round"] is None][0]
        assert rel_close(stats[w][0], unhex(pica["pi"]), TOL)
        hf = unhex([r for r in case["hfst"] if r["L"] == case["L"] and r["round"] is None][0]["res"])
        for col, key in ((2, "pi_a"), (3, "pi_b"), (4, "pi_xy"), (5, "dxy"), (6, "da"), (7, "fst")):
            assert rel_close(stats[w][col], hf[key], TOL), key
    batch.close()


SHAPES = [(1, 5, 1), (2, 1, 2), (3, 33, 3), (17, 64, 4), (40, 300, 5), (127, 100, 6), (128, 129, 7), (129, 257, 8),
          (257, 200, 9), (300, 1000, 10), (466, 1024, 11), (513, 70, 12)]


@pytest.mark.parametrize("algo", ALGOS)
def test_pairwise_random_ragged(ctx, algo):
    """Ragged batch incl. heavy nodes (len >= 255 up to 2e5), zero-length nodes, identical / empty rows."""
    from impop_b200.engine import WindowBatch
    wins, raw = [], []
    for n, m, seed in SHAPES:
        x, node_len = _random_window(n, m, seed)
        bits = similarity.pack_bits(x)
        ia, ib = list(range(0, n, 3)), list(range(1, n, 3))
        sub = list(range(0, n, 2)) if n > 3 else list(range(n))
        lab = _labels(n, ia, ib, subset=sub, seg=sub)
        wins.append((bits, node_len, lab, 1000 + seed))
        raw.append((x, node_len))
    batch = WindowBatch.from_windows(ctx, wins)
    for w, (bits, node_len, lab, L) in enumerate(wins):
        n, m = SHAPES[w][0], SHAPES[w][1]
        A0, I0, pi0 = clib.window_pairwise(bits, m, node_len)
        I, A, pi = batch.pairwise(w, algo)
        ctx.check()
        assert (A.cpu().numpy() == A0).all(), f"A mismatch window {w}"
        assert (I.cpu().numpy() == I0).all(), f"I mismatch window {w} shape {SHAPES[w]}"
        assert (pi.cpu().numpy() == pi0).all(), f"pi mismatch window {w}"
    stats, counts = batch.stats(algo)
    ctx.check()
    want_s, want_c = clib.batch_stats(batch.n, batch.m, batch.pitch_words, batch.x_off, batch.len_off, batch.lab_off,
                                      batch.length, batch.x.cpu().numpy().view(np.uint32),
                                      batch.node_len.cpu().numpy().view(np.uint32), batch.labels.cpu().numpy(), 4)
    assert (counts.cpu().numpy() == want_c).all()
    got = stats.cpu().numpy()
    for w in range(len(wins)):
        _assert_stats(got[w], want_s[w], f"window {w} shape {SHAPES[w]}")
    batch.close()


@pytest.mark.parametrize("algo", ALGOS)
def test_very_heavy_nodes(ctx, algo):
    """Node lengths up to 2^24 (three byte planes in the north star's wording)."""
    from impop_b200.engine import WindowBatch
    rng = np.random.default_rng(99)
    n, m = 70, 90
    x = (rng.random((n, m)) < 0.5).astype(np.uint8)
    node_len = rng.integers(1, 1 << 24, size=m).astype(np.uint32)
    node_len[::7] = 255
    node_len[1::7] = 254
    node_len[2::7] = 65025
    bits = similarity.pack_bits(x)
    batch = WindowBatch.from_windows(ctx, [(bits, node_len, _labels(n, range(0, 30), range(30, 70)), 0)])
    A0, I0, pi0 = clib.window_pairwise(bits, m, node_len)
    I, A, pi = batch.pairwise(0, algo)
    ctx.check()
    assert (A.cpu().numpy() == A0).all() and (I.cpu().numpy() == I0).all() and (pi.cpu().numpy() == pi0).all()
    batch.close()


def test_hprc_shaped_batch_both_algos_agree(ctx):
    """SURVEY 8(d) generator, config-2 shape (466 haplotypes, 50 kb): tcgen05 == SIMT == oracle."""
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(466, 50000, 6, seed=0xB200 + 2)
    pops = ws.pops
    lab = _labels(466, np.nonzero(pops == 0)[0], np.nonzero(pops == 2)[0])
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    s_tc, c_tc = batch.stats(0)
    s_si, c_si = batch.stats(1)
    ctx.check()
    want_s, want_c = clib.batch_stats(batch.n, batch.m, batch.pitch_words, batch.x_off, batch.len_off, batch.lab_off,
                                      batch.length, ws.x_bits, ws.node_len, lab, 4)
    assert (c_tc.cpu().numpy() == want_c).all() and (c_si.cpu().numpy() == want_c).all()
    for w in range(ws.windows):
        _assert_stats(s_tc.cpu().numpy()[w], want_s[w], f"tc window {w}")
        _assert_stats(s_si.cpu().numpy()[w], want_s[w], f"simt window {w}")
    # raw sums are fixed-order reductions: each algorithm is run-to-run reproducible
    s_tc2, _ = batch.stats(0)
    ctx.check()
    assert torch.equal(s_tc.nan_to_num(7.0), s_tc2.nan_to_num(7.0))
    batch.close()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_grid_split_matches_single(ctx, world):
    """work items dealt round-robin to `world` ranks, partial sums added in rank order."""
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(600, 20000, 3, seed=5, n_sites_override=60)
    lab = _labels(600, range(0, 200), range(200, 450))
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    parts = torch.stack([batch.window_sums(r, world) for r in range(world)]).contiguous()
    stats, counts = batch.finalize(parts)
    ctx.check()
    want_s, want_c = clib.batch_stats(batch.n, batch.m, batch.pitch_words, batch.x_off, batch.len_off, batch.lab_off,
                                      batch.length, ws.x_bits, ws.node_len, lab, 4)
    assert (counts.cpu().numpy() == want_c).all()
    for w in range(ws.windows):
        _assert_stats(stats.cpu().numpy()[w], want_s[w], f"split window {w}")
    batch.close()


def test_properties_at_scale(ctx):
    """Size-independent properties on a batch too large for the oracle to cover pair by pair:
    (1) permuting haplotypes leaves every statistic unchanged to 1e-12 and S exactly;
    (2) duplicating a window gives bit-identical rows; (3) pi_xy == (pi_a + pi_b) / 2 exactly."""
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(466, 50000, 64, seed=77)
    lab = _labels(466, np.nonzero(ws.pops == 0)[0], np.nonzero(ws.pops == 2)[0])
    perm = np.random.default_rng(1).permutation(466)
    xb = np.concatenate([ws.x_bits, ws.x_bits[:, perm, :], ws.x_bits[:1]])
    nl = np.concatenate([ws.node_len, ws.node_len, ws.node_len[:1]])
    labs = np.concatenate([np.tile(lab, (64, 1)), np.tile(lab[perm], (64, 1)), lab[None]])
    batch = WindowBatch.from_uniform(ctx, xb, nl, labs, ws.length)
    stats, counts = batch.stats(0)
    ctx.check()
    s, c = stats.cpu().numpy(), counts.cpu().numpy()
    assert (c[:64] == c[64:128]).all()
    for w in range(64):
        _assert_stats(s[64 + w], s[w], f"permuted window {w}")
    assert (np.nan_to_num(s[128], nan=7.0) == np.nan_to_num(s[0], nan=7.0)).all()
    L = float(ws.length)
    assert np.allclose(s[:, 4] * L, 0.5 * (s[:, 2] * L + s[:, 3] * L), rtol=1e-15, atol=0)
    # spot-check a few windows pair by pair against the oracle
    for w in (0, 31, 63):
        want_s, want_c = clib.window_stats(ws.x_bits[w], ws.m_pad, ws.node_len[w], lab, ws.length)
        assert (c[w] == want_c).all()
        _assert_stats(s[w], want_s, f"oracle window {w}")
    batch.close()


def test_edge_cases(ctx):
    """Empty batch, n in {0, 1}, m == 0, no population members, S == 0."""
    from impop_b200.engine import WindowBatch
    z = np.zeros((0, 4), dtype=np.uint32)
    one = np.array([[0b1011, 0, 0, 0]], dtype=np.uint32)
    two = np.array([[0b1011, 0, 0, 0], [0b1011, 0, 0, 0]], dtype=np.uint32)
    wins = [
        (z, np.array([3, 4, 5, 6], dtype=np.uint32), np.zeros(0, dtype=np.uint8), 100),                 # n = 0
        (one, np.array([3, 4, 5, 6], dtype=np.uint32), np.array([15], dtype=np.uint8), 100),             # n = 1
        (two, np.array([3, 4, 5, 6], dtype=np.uint32), np.array([9, 9], dtype=np.uint8), 100),           # identical, no A/B
        (np.zeros((3, 4), dtype=np.uint32), np.zeros(0, dtype=np.uint32), np.array([11, 13, 9], dtype=np.uint8), 0),  # m = 0
    ]
    batch = WindowBatch.from_windows(ctx, wins)
    for algo in ALGOS:
        stats, counts = batch.stats(algo)
        ctx.check()
        s, c = stats.cpu().numpy(), counts.cpu().numpy()
        for w, (bits, nl, lab, L) in enumerate(wins):
            want_s, want_c = clib.window_stats(bits, len(nl), nl, lab, L)
            assert (c[w] == want_c).all(), (algo, w)
            _assert_stats(s[w], want_s, f"edge {w} algo {algo}")
    batch.close()
    empty = WindowBatch.from_windows(ctx, [])
    st, ct = empty.stats(0)
    ctx.check()
    assert st.shape == (0, 20) and ct.shape == (0, 8)
    empty.close()


def test_division_selftest(ctx):
    """The epilogue's range-restricted division (no slow-path checks) is bit-identical to __ddiv_rn on
    2^26 pseudo-random (I, A_i, A_j) triples incl. I == 0, I == min(A), tiny and huge paths."""
    assert ctx.selftest_division(1 << 26, seed=7) == 0
    assert ctx.selftest_division(1 << 24, seed=12345) == 0


def test_many_small_windows_and_wide_rows(ctx):
    """Ragged batch that exercises every row-block / column-split shape of the tile scheduler
    (n from 1 to 700) with few nodes, plus one window with m > 2048 (two passes of the length table)."""
    from impop_b200.engine import WindowBatch
    rng = np.random.default_rng(21)
    wins = []
    for n in (1, 2, 15, 16, 17, 31, 33, 64, 100, 128, 129, 130, 200, 255, 256, 257, 300, 383, 384, 385, 400, 512, 513, 640, 700):
        m = int(rng.integers(1, 90))
        x = (rng.random((n, m)) < 0.5).astype(np.uint8)
        nl = rng.integers(0, 300, size=m).astype(np.uint32)
        lab = _labels(n, range(0, n, 2), range(1, n, 2))
        wins.append((similarity.pack_bits(x), nl, lab, 777))
    x = (rng.random((150, 3000)) < 0.3).astype(np.uint8)
    nl = rng.integers(0, 500, size=3000).astype(np.uint32)
    wins.append((similarity.pack_bits(x), nl, _labels(150, range(0, 70), range(70, 150)), 9999))
    batch = WindowBatch.from_windows(ctx, wins)
    want_s, want_c = clib.batch_stats(batch.n, batch.m, batch.pitch_words, batch.x_off, batch.len_off, batch.lab_off,
                                      batch.length, batch.x.cpu().numpy().view(np.uint32),
                                      batch.node_len.cpu().numpy().view(np.uint32), batch.labels.cpu().numpy(), 4)
    for algo in ALGOS:
        stats, counts = batch.stats(algo)
        ctx.check()
        assert (counts.cpu().numpy() == want_c).all()
        got = stats.cpu().numpy()
        for w in range(len(wins)):
            _assert_stats(got[w], want_s[w], f"algo {algo} window {w} n={wins[w][0].shape[0]}")
    w = len(wins) - 1
    A0, I0, pi0 = clib.window_pairwise(wins[w][0], 3000, wins[w][1])
    I, A, pi = batch.pairwise(w, 0)
    ctx.check()
    assert (A.cpu().numpy() == A0).all() and (I.cpu().numpy() == I0).all() and (pi.cpu().numpy() == pi0).all()
    batch.close()


def test_range_error_is_reported(ctx):
    """sum(node_len) >= 2^31 is refused loudly (exactness bound of the int32 accumulators)."""
    from impop_b200 import _native
    from impop_b200.engine import WindowBatch
    bits = np.full((4, 4), 0xFFFFFFFF, dtype=np.uint32)
    node_len = np.full(128, 1 << 25, dtype=np.uint32)
    batch = WindowBatch.from_windows(ctx, [(bits, node_len, np.full(4, 15, dtype=np.uint8), 0)])
    batch.stats(1)
    with pytest.raises(_native.NativeError) as err:
        ctx.check()
    assert err.value.code == -4
    batch.close()


def test_wrapper_tsv_rows(ctx, tmp_path):
    """Matrix mode end to end: batch -> stats -> the wrappers' TSV columns (run_h-fst.sh:148, run_tajd.sh:101)."""
    import io
    from impop_b200 import windows
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(90, 100000, 3, seed=0xB200 + 1, n_sites_override=40)
    lab = _labels(90, range(0, 30), range(30, 70))
    batch = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    stats, counts = batch.stats(0)
    ctx.check()
    st, ct = stats.cpu().numpy(), counts.cpu().numpy()
    regions = [windows.region_name("chr2", 100000 * k, 100000 * (k + 1)) for k in range(3)]
    buf = io.StringIO()
    windows.write_tsv(buf, "fst", windows.fst_rows(regions, [ws.length] * 3, st))
    lines = buf.getvalue().splitlines()
    assert lines[0].split("\t") == ["REGION", "LENGTH", "FST", "PI_A", "PI_B", "PI_XY", "DXY", "DA"]
    want_s, want_c = clib.window_stats(ws.x_bits[1], ws.m_pad, ws.node_len[1], lab, ws.length)
    f = lines[2].split("\t")
    assert f[0] == "CHM13#0#chr2:100000-200000" and f[1] == "100000"
    assert f[2:] == [f"{want_s[k]:.8f}" for k in (7, 2, 3, 4, 5, 6)]
    buf = io.StringIO()
    windows.write_tsv(buf, "tajd", windows.tajd_rows(regions, [ws.length] * 3, st, ct))
    t = buf.getvalue().splitlines()[2].split("\t")
    assert t[2] == "90" and t[3] == str(int(want_c[7])) and t[4] == f"{want_s[1]:.8f}"
    assert t[5] == ("NA" if np.isnan(want_s[9]) else repr(float(st[1][9])))
    pi = list(windows.pi_rows(regions, [ws.length] * 3, st))
    assert pi[0][-1] == f"{want_s_first(ws, lab):.8f} (sequence length: 100000)"
    batch.close()


def want_s_first(ws, lab):
    return clib.window_stats(ws.x_bits[0], ws.m_pad, ws.node_len[0], lab, ws.length)[0][1]
