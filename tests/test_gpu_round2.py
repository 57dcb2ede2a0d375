"""Round-2 additions on the device: exact decimal rounding, ingest-time column compaction through the
fused kernels, the torch-free command lines (lite context), stream-aware scratch pool."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from impop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


HARD = [0.5, 1.5, 2.5, 0.125, 0.375, 0.625, 0.0625, 0.15, 0.25, 0.35, 0.45, 2.675, 1.005, 0.285, 1.115, 8.345, 0.00005, 0.000015,
        0.99999, 0.999995, 0.9999949999999999, 0.9999950000000001, 1e-9, 5e-6, 4.9999999999999996e-06, 123456.789, 1e15, 4503599627370497.0,
        1e22, 1e300, 0.0, -0.0, -0.5, -1.5, -2.675, -0.125, float("nan"), float("inf"), -float("inf"),
        0.49999999999999994, 0.5000000000000001, 5e-324, 2.2250738585072014e-308]


@pytest.mark.parametrize("digits", [0, 1, 2, 3, 5, 8, 12, 17, 22])
def test_round_decimal_matches_cpython(ctx, digits):
    import torch
    rng = np.random.default_rng(100 + digits)
    vals = np.concatenate([
        np.array(HARD, dtype=np.float64),
        rng.random(400_000),                                             # identities
        1.0 - rng.random(200_000) * 1e-3,                                # near 1, as real identities are
        np.round(rng.random(100_000), min(digits + 1, 15)) ,             # decimal ties and near-ties at the next digit
        (rng.integers(0, 10 ** min(digits + 1, 15), 100_000) * 2 + 1) / (2.0 * 10.0 ** min(digits, 15)),   # k + 1/2 patterns
        rng.standard_normal(50_000) * 10.0 ** rng.integers(-8, 12, 50_000),
    ])
    dev = torch.from_numpy(vals.copy()).to(ctx.torch_device)
    ctx.round_decimal(dev, digits)
    ctx.check()
    got = dev.cpu().numpy()
    want = np.array([v if (v != v or math.isinf(v)) else round(v, digits) for v in vals.tolist()], dtype=np.float64)
    same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want)) | ((got == 0) & (want == 0))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, [(float(vals[k]), float(got[k]), float(want[k])) for k in bad[:5]]


def test_round_digits_through_the_table_path(ctx):
    """pica2 -r / h-fst -r: the table's device copy is rounded by the kernel; same values as a host round() per element."""
    from impop_b200.tables import SimilarityTable
    rng = np.random.default_rng(5)
    n = 60
    m = 1.0 - rng.random((n, n)) * 1e-2
    m = np.triu(m, 1) + np.triu(m, 1).T
    m[3, 7] = m[7, 3] = np.nan
    tab = SimilarityTable([f"s{i:03d}" for i in range(n)], m)
    for digits in (3, 5):
        dev = tab.device(ctx, digits).cpu().numpy()
        want = tab.rounded(digits)
        assert np.array_equal(np.isnan(dev), np.isnan(want))
        assert np.array_equal(dev[~np.isnan(dev)], want[~np.isnan(want)])


def test_compacted_batch_gives_the_same_rows(ctx):
    """Ingest-time column compaction (impop_compact_scan / _fill), plain and affine form (bubbles merged, constants in C,
    weights spread over copies), leaves every count and statistic of the fused path as it was: ragged windows with
    constant, empty, zero-length, heavy, identical and complementary columns, both algorithms."""
    from impop_b200 import ingest
    from impop_b200.engine import ALGO_SIMT, ALGO_TCGEN05, WindowBatch
    from oracle import similarity
    rng = np.random.default_rng(77)
    wins, cwins, awins = [], [], []
    for n, m, heavy in ((37, 70, False), (130, 300, True), (466, 1009, False), (200, 1500, True), (5, 3, False)):
        x = (rng.random((n, m)) < rng.random(m)[None, :]).astype(np.uint8)
        kind = rng.random(m)
        x[:, kind < 0.3] = 1
        x[:, (kind >= 0.3) & (kind < 0.4)] = 0
        for k in range(1, m - 1, 3):                         # bubbles (complementary columns) and variants in perfect linkage
            if kind[k] >= 0.4 and kind[k] < 0.7:
                x[:, k + 1] = 1 - x[:, k]
            elif kind[k] >= 0.95:
                x[:, k + 1] = x[:, k]
        nl = rng.integers(1, 100000 if heavy else 60, size=m).astype(np.uint32)
        nl[rng.random(m) < 0.5] = 1
        nl[rng.random(m) < 0.1] = 0
        lab = np.full(n, 9, dtype=np.uint8)
        lab[: n // 3] |= 2
        lab[n // 3: n // 2] |= 4
        lab[rng.random(n) < 0.1] &= 0xF6                     # a few rows outside SUBSET / SEG
        bits = similarity.pack_bits(x)
        wins.append((bits, nl, lab, 1234))
        win = ingest.GraphWindow([f"h{i}" for i in range(n)], bits, nl)
        g = ingest.compact_window(win, pairs=False)
        assert g.m < m
        cwins.append((g.x_bits, g.node_len, lab, 1234))
        h = ingest.compact_window(win)
        assert int((h.col_mult > 0).sum()) < g.m or n < 10      # fewer columns before weights are spread over copies
        awins.append((h.x_bits, h.node_len, lab, 1234, h.row_adj, h.win_const, h.col_mult))
    a = WindowBatch.from_windows(ctx, wins)
    for other in (cwins, awins, [awins[0], wins[1], awins[2], wins[3], awins[4]]):      # (the last: affine and plain windows in one batch)
        b = WindowBatch.from_windows(ctx, other)
        for algo in (ALGO_TCGEN05, ALGO_SIMT):
            sa, ca = a.stats(algo)
            sb, cb = b.stats(algo)
            ctx.check()
            assert np.array_equal(ca.cpu().numpy(), cb.cpu().numpy())
            sa, sb = sa.cpu().numpy()[:, :19], sb.cpu().numpy()[:, :19]       # column 19 (variant sites) depends on the node order
            assert np.array_equal(np.isnan(sa), np.isnan(sb))
            ok = np.isnan(sa) | (np.abs(sa - sb) <= 1e-13 * np.maximum(np.abs(sa), np.abs(sb)))
            assert ok.all(), np.argwhere(~ok)[:5]
            for w in (1, 4):
                I0, A0, p0 = a.pairwise(w, algo)
                I1, A1, p1 = b.pairwise(w, algo)
                ctx.check()
                assert np.array_equal(I0.cpu().numpy(), I1.cpu().numpy()) and np.array_equal(A0.cpu().numpy(), A1.cpu().numpy())
                assert np.array_equal(p0.cpu().numpy(), p1.cpu().numpy())            # pi_ij bit for bit
        b.close()
    a.close()


def test_command_lines_run_without_torch(tmp_path):
    """TSV mode: scripts/pica2.py, h-fst.py, af.py and tj_d.py use the lite context (plain device buffers over the C ABI);
    the wrappers start them once per window, so the torch import must never happen there."""
    from impop_b200 import synth
    from oracle import similarity
    ws = synth.make_windows(24, 20000, 1, seed=3)
    res = similarity.pairwise(ws.dense(0), ws.node_len[0])
    names = synth.haplotype_names(24, "chr2", 1000, 21000)
    tsv = tmp_path / "w.sim.tsv"
    similarity.write_similarity_tsv(str(tsv), names, res)
    (tmp_path / "a.txt").write_text("\n".join(synth.assembly_names(range(0, 8))) + "\n")
    (tmp_path / "b.txt").write_text("\n".join(synth.assembly_names(range(8, 16))) + "\n")
    for mod, argv in (("pica2", [str(tsv), "-t", "1.0", "-l", "20000", "-r", "5", "-d", str(tmp_path)]),
                      ("hfst", [str(tsv), "-a", str(tmp_path / "a.txt"), "-b", str(tmp_path / "b.txt"), "-l", "20000", "-r", "4", "-d", str(tmp_path)]),
                      ("af", ["--input", str(tsv), "--threshold", "0.999", "--output", str(tmp_path / "af.tsv")]),
                      ("tj_d", ["-n", "446", "-p", "0.59146123", "-S", "20"])):
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); sys.argv = ['x'] + {argv!r}\n"
                f"from impop_b200.{mod} import main\n"
                "rc = main()\n"
                "assert 'torch' not in sys.modules, 'torch was imported'\n"
                "sys.exit(rc or 0)\n")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        assert r.returncode == 0, (mod, r.stdout[-500:], r.stderr[-1500:])
        if mod == "tj_d":
            assert r.stdout.strip() == "Tajima's D: -1.992648227415639"


def test_scratch_pool_across_streams(ctx):
    """A batch destroyed while its kernels are in flight: its blocks are not handed to another stream's batch before the
    first stream's work has finished (event-guarded pool) -- results of interleaved create / stats / destroy on two streams
    equal the serial ones."""
    import torch
    from impop_b200 import synth
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(200, 20000, 24, seed=11)
    lab = np.full(200, 9, dtype=np.uint8)
    ref = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, 20000)
    want_s, want_c = ref.stats()
    ctx.check()
    want_s, want_c = want_s.cpu().numpy(), want_c.cpu().numpy()
    ref.close()
    xd = torch.from_numpy(ws.x_bits.view(np.int32)).to(ctx.torch_device)
    ld = torch.from_numpy(ws.node_len.view(np.int32)).to(ctx.torch_device)
    labd = torch.from_numpy(lab).to(ctx.torch_device)
    streams = [torch.cuda.Stream(device=ctx.torch_device) for _ in range(2)]
    torch.cuda.synchronize()
    outs = []
    for rep in range(6):
        st = streams[rep % 2]
        with torch.cuda.stream(st):
            b = WindowBatch.from_uniform(ctx, xd, ld, labd, 20000, stream=st)
            s, c = b.stats(stream=st)
            b.close()                                     # kernels still in flight
            outs.append((s, c))
    torch.cuda.synchronize()
    ctx.check()
    for s, c in outs:
        assert np.array_equal(c.cpu().numpy(), want_c)
        got = s.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want_s))
        assert (np.isnan(got) | (np.abs(got - want_s) <= 1e-12 * np.abs(want_s))).all()


def test_variant_sites_column(ctx):
    """IMPOP_ST_S_BUBBLES: device count (node order of the batch, SEG rows) against the restatement; after ingest-time
    compaction the count taken on the original order travels with the batch."""
    from impop_b200 import ingest, synth
    from impop_b200._native import ST
    from impop_b200.engine import WindowBatch
    from oracle import clib, similarity
    rng = np.random.default_rng(21)
    wins, want = [], []
    for n, m in ((37, 70), (130, 300), (466, 1009), (64, 2100), (9, 33)):
        x = (rng.random((n, m)) < rng.random(m)[None, :]).astype(np.uint8)
        kind = rng.random(m)
        x[:, kind < 0.35] = 1
        x[:, (kind >= 0.35) & (kind < 0.45)] = 0
        nl = rng.integers(0, 5, size=m).astype(np.uint32)
        lab = np.full(n, 1, dtype=np.uint8)
        seg_rows = np.flatnonzero(rng.random(n) < 0.7)
        lab[seg_rows] |= 8
        wins.append((similarity.pack_bits(x), nl, lab, 1000))
        want.append(similarity.site_runs(x, nl, seg_rows))
    b = WindowBatch.from_windows(ctx, wins)
    st, ct = b.stats()
    ctx.check()
    st = st.cpu().numpy()
    assert st[:, ST["S_bubbles"]].astype(int).tolist() == want
    ws_, wc_ = clib.batch_stats(b.n, b.m, b.pitch_words, b.x_off, b.len_off, b.lab_off, b.length, b.x.cpu().numpy().view(np.uint32),
                                b.node_len.cpu().numpy().view(np.uint32), b.labels.cpu().numpy(), 2)
    assert ws_[:, 19].astype(int).tolist() == want                          # the C oracle states the same definition
    b.close()
    ws = synth.make_windows(60, 20000, 4, seed=3)
    c = ingest.compact_uniform(ws.x_bits, ws.node_len)
    b = WindowBatch.from_uniform(ctx, c.x, c.node_len, np.full(60, 9, dtype=np.uint8), 20000, **c.batch_kwargs())
    st = b.stats()[0].cpu().numpy()
    ctx.check()
    assert st[:, ST["S_bubbles"]].astype(int).tolist() == [similarity.site_runs(ws.dense(w), ws.node_len[w]) for w in range(4)]
    assert (2 * st[:, ST["S_bubbles"]] == st[:, ST["S"]]).all()              # bi-allelic bubbles: two segregating nodes per site
    b.close()


def test_repitch_rows(ctx):
    """impop_repitch_rows: tight rows (as stored / transferred) -> 16-byte-multiple rows, zero padded; ragged windows."""
    import torch
    rng = np.random.default_rng(11)
    rows = np.array([5, 0, 466, 37, 1], dtype=np.int32)
    sp = np.array([9, 3, 11, 1, 4], dtype=np.int32)
    dp = np.array([12, 4, 12, 4, 4], dtype=np.int32)
    so = np.concatenate([[0], np.cumsum(rows.astype(np.int64) * sp)])
    do = np.concatenate([[0], np.cumsum(rows.astype(np.int64) * dp)])
    src = rng.integers(0, 2 ** 32, size=int(so[-1]), dtype=np.uint64).astype(np.uint32)
    dsrc = torch.from_numpy(src.view(np.int32)).to(ctx.torch_device)
    ddst = torch.full((int(do[-1]),), -1, dtype=torch.int32, device=ctx.torch_device)
    ctx.repitch_rows(dsrc, ddst, rows, sp, dp, so[:-1], do[:-1])
    ctx.check()
    got = ddst.cpu().numpy().view(np.uint32)
    for w in range(len(rows)):
        a = src[so[w]:so[w + 1]].reshape(rows[w], sp[w])
        b = got[do[w]:do[w + 1]].reshape(rows[w], dp[w])
        assert np.array_equal(b[:, :sp[w]], a) and not b[:, sp[w]:].any(), w
    with pytest.raises(Exception):
        ctx.repitch_rows(dsrc, ddst, rows, sp, dp - 1, so[:-1], do[:-1])      # destination pitch not a multiple of 4 words


def test_heavy_entry_counts_from_the_caller(ctx):
    """impop_batch_desc_t.heavy_entries_host: with the caller's counts batch set-up skips its pass over the node lengths; the
    results are the same, and counts that are too small are caught on the device."""
    from impop_b200 import ingest, synth
    from impop_b200.engine import WindowBatch
    from impop_b200._native import NativeError
    ws = synth.make_windows(40, 20000, 3, seed=21)
    nl = ws.node_len.copy()
    nl[:, 5] = 70000                                          # nodes of >= 255 bp: heavy entries
    nl[1, 9] = 300
    lab = np.full(40, 9, dtype=np.uint8)
    ref = WindowBatch.from_uniform(ctx, ws.x_bits, nl, lab, 20000)
    s0, c0 = ref.stats()
    ctx.check()
    he = ingest.heavy_entries(nl)
    assert he.tolist() == [int(((l.astype(np.int64) // 255 + 254) // 255).sum()) for l in nl] and he.min() >= 2
    b = WindowBatch.from_uniform(ctx, ws.x_bits, nl, lab, 20000, heavy_entries=he)
    s1, c1 = b.stats()
    ctx.check()
    assert np.array_equal(c0.cpu().numpy(), c1.cpu().numpy())
    a, bb = s0.cpu().numpy(), s1.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(bb)) and np.array_equal(a[~np.isnan(a)], bb[~np.isnan(bb)])
    bad = WindowBatch.from_uniform(ctx, ws.x_bits, nl, lab, 20000, heavy_entries=np.zeros(3, dtype=np.int32))
    bad.stats()
    with pytest.raises(NativeError):
        ctx.check()
    ref.close(); b.close(); bad.close()


def test_affine_window_without_columns(ctx):
    """A window whose nodes every haplotype visits (and one that loses all but a bubble): the affine form has no column
    left -- everything is in the window constant -- and still gives the plain batch's rows."""
    from impop_b200 import ingest
    from impop_b200.engine import ALGO_SIMT, ALGO_TCGEN05, WindowBatch
    from oracle import similarity
    lab = np.array([9, 11, 13, 9, 11, 13], dtype=np.uint8)
    xs = [np.ones((6, 4), np.uint8),
          np.array([[1, 1, 0, 1], [1, 0, 1, 1], [1, 1, 0, 1], [1, 0, 1, 1], [1, 1, 0, 1], [1, 1, 0, 1]], np.uint8)]
    nls = [np.array([3, 5, 7, 0], np.uint32), np.array([100, 1, 1, 50], np.uint32)]
    plain, aff = [], []
    for x, nl in zip(xs, nls):
        bits = similarity.pack_bits(x)
        plain.append((bits, nl, lab, 200))
        g = ingest.compact_window(ingest.GraphWindow([f"h{i}" for i in range(6)], bits, nl))
        aff.append((g.x_bits, g.node_len, lab, 200, g.row_adj, g.win_const, g.col_mult))
    assert aff[0][1].shape[0] == 0 and aff[0][5] == 15 and aff[1][1].tolist() == [2] and aff[1][5] == 151
    a, b = WindowBatch.from_windows(ctx, plain), WindowBatch.from_windows(ctx, aff)
    for algo in (ALGO_TCGEN05, ALGO_SIMT):
        sa, ca = a.stats(algo)
        sb, cb = b.stats(algo)
        ctx.check()
        assert np.array_equal(ca.cpu().numpy(), cb.cpu().numpy())
        sa, sb = sa.cpu().numpy()[:, :19], sb.cpu().numpy()[:, :19]
        assert np.array_equal(np.isnan(sa), np.isnan(sb)) and np.array_equal(sa[~np.isnan(sa)], sb[~np.isnan(sb)])
        for w in range(2):
            I0, A0, p0 = a.pairwise(w, algo)
            I1, A1, p1 = b.pairwise(w, algo)
            ctx.check()
            assert np.array_equal(I0.cpu().numpy(), I1.cpu().numpy()) and np.array_equal(A0.cpu().numpy(), A1.cpu().numpy())
            assert np.array_equal(p0.cpu().numpy(), p1.cpu().numpy())
    a.close(); b.close()
