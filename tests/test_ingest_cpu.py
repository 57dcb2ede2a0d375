"""Host-side ingest: the C reader of libimpop_b200 against the pure-Python restatement
(oracle/gfa.py) on GFA v1 text with P and W lines, '*' sequences, repeated visits, CRLF, foreign record types."""
import io

import numpy as np
import pytest

from impop_b200 import ingest, synth
from impop_b200._native import NativeError
from oracle import gfa as ogfa
from oracle import similarity

HAND = ("H\tVN:Z:1.0\n"
        "S\ts1\tACGT\n"
        "S\t2\t*\tLN:i:300\n"
        "S\tnode_3\tA\tRC:i:5\n"
        "S\t44\t*\n"
        "L\ts1\t+\t2\t+\t0M\n"
        "P\tHG00097#1#CM094061.1:100-200\ts1+,2-,node_3+\t*\n"
        "P\tHG00097#2#CM094062.1:100-200\ts1+,node_3+,s1+,s1-\t4M,1M\n"
        "W\tHG002\t1\tchr2\t10\t90\t>2<44>2\n"
        "W\tCHM13\t0\tchr2\t*\t*\t>s1\n"
        "P\tempty#1#ctg\t*\t*\r\n"
        "# trailing comment\n")


def check(text):
    names, x, counts, node_len = ogfa.parse(text)
    win = ingest.parse_gfa(text, want_counts=True)
    assert win.names == names
    assert np.array_equal(win.node_len.astype(np.int64), node_len)
    assert np.array_equal(similarity.unpack_bits(win.x_bits, len(node_len)), x)
    assert np.array_equal(win.counts.astype(np.int64), np.minimum(counts, 65535))
    assert win.x_bits.shape[1] % 4 == 0
    return win


def test_hand_written_gfa():
    win = check(HAND)
    assert win.names[2] == "HG002#1#chr2:10-90" and win.names[3] == "CHM13#0#chr2"
    assert win.node_len.tolist() == [4, 300, 1, 0]
    assert win.counts[1].tolist() == [3, 0, 1, 0]


@pytest.mark.parametrize("walks", [False, True])
def test_synthetic_window_round_trip(walks):
    ws = synth.make_windows(40, 20000, 1, seed=77)
    names = synth.haplotype_names(40, "chr2", 1000, 21000)
    buf = io.StringIO()
    ingest.write_gfa(buf, names, ws.dense(0)[:, :ws.m], ws.node_len[0, :ws.m], walks=walks)
    win = check(buf.getvalue())
    assert win.names == names and win.m == ws.m
    assert np.array_equal(similarity.unpack_bits(win.x_bits, win.m), ws.dense(0)[:, :ws.m])
    assert np.array_equal(win.node_len, ws.node_len[0, :ws.m])


def test_empty_and_errors():
    win = ingest.parse_gfa("H\tVN:Z:1.0\n")
    assert win.n == 0 and win.m == 0
    with pytest.raises(NativeError):
        ingest.parse_gfa("S\t1\tA\nP\tp\t1+,9+\t*\n")          # step over an undefined segment
    with pytest.raises(NativeError):
        ingest.parse_gfa("S\t1\tA\nS\t1\tC\n")                  # duplicate segment name
    with pytest.raises(NativeError):
        ingest.parse_gfa("S\t1\tA\nP\tp\t1\t*\n")               # step without orientation


def test_batch_container(tmp_path):
    ws = synth.make_windows(12, 5000, 3, seed=5)
    wins = []
    for w in range(3):
        names = synth.haplotype_names(12, "chr2", w * 5000, (w + 1) * 5000)
        wins.append(ingest.GraphWindow(names, ws.x_bits[w].copy(), ws.node_len[w, :ws.m].copy(),
                                       None, f"CHM13#0#chr2:{w * 5000}-{(w + 1) * 5000}", 5000))
    path = tmp_path / "batch.npz"
    ingest.save_batch(path, wins)
    back = ingest.load_batch(path)
    assert len(back) == 3
    for a, b in zip(wins, back):
        assert a.names == b.names and a.region == b.region and a.length == b.length
        assert np.array_equal(a.x_bits, b.x_bits) and np.array_equal(a.node_len, b.node_len)


def test_multiset_expansion_is_the_min_count_intersection():
    """f-4: set intersection over copy nodes == sum len * min(count_i, count_j); path length == sum len * count."""
    rng = np.random.default_rng(9)
    n, m = 14, 37
    counts = rng.integers(0, 4, size=(n, m))
    counts[:, 5] = 0                                             # a node nobody visits
    counts[3] = 0                                                # an empty path
    node_len = rng.integers(0, 400, size=m)
    lines = ["H\tVN:Z:1.0"] + [f"S\t{k + 1}\t*\tLN:i:{node_len[k]}" for k in range(m)]
    for i in range(n):
        steps = [f"{k + 1}+" for k in range(m) for _ in range(counts[i, k])]
        rng.shuffle(steps)
        lines.append(f"P\tS{i:03d}#1#ctg:0-100\t" + (",".join(steps) if steps else "*") + "\t*")
    win = ingest.parse_gfa("\n".join(lines) + "\n", want_counts=True)
    assert np.array_equal(win.counts.astype(np.int64), counts)
    ex = ingest.multiset_expand(win)
    x = similarity.unpack_bits(ex.x_bits, ex.m).astype(np.int64)
    inter = (x * ex.node_len.astype(np.int64)[None, :]) @ x.T
    want = np.einsum("k,ijk->ij", node_len.astype(np.int64), np.minimum(counts[:, None, :], counts[None, :, :]))
    assert np.array_equal(inter, want)
    assert np.array_equal(np.diag(inter), counts @ node_len)
    assert ex.m == int(np.maximum(counts.max(axis=0), 1).sum())


def test_read_gfa_many_matches_one_by_one(tmp_path):
    """The threaded reader of a chromosome's window graphs returns the windows in input order, identical to reading
    them one by one; a malformed file raises as it does in read_gfa."""
    from impop_b200._native import NativeError
    ws = synth.make_windows(40, 3000, 3, seed=21)
    names = synth.haplotype_names(40, "chr2", 0, 3000)
    items = []
    for w in range(3):
        for rep in range(3):
            p = tmp_path / f"w{w}_{rep}.gfa"
            with open(p, "w") as fh:
                ingest.write_gfa(fh, names, ws.dense(w)[:, :ws.m], ws.node_len[w][:ws.m], walks=bool(rep & 1))
            items.append((f"chr2:{3000 * w}-{3000 * (w + 1)}", str(p), 3000))
    many = ingest.read_gfa_many(items, threads=4)
    for it, g in zip(items, many):
        one = ingest.read_gfa(it[1], region=it[0], length=it[2])
        assert g.region == it[0] and g.length == 3000 and g.names == one.names
        assert np.array_equal(g.x_bits, one.x_bits) and np.array_equal(g.node_len, one.node_len)
    assert [g.region for g in ingest.read_gfa_many(items, threads=1)] == [it[0] for it in items]
    bad = tmp_path / "bad.gfa"
    bad.write_text("S\t1\tACGT\nP\tp1\t1+,2+\t*\n")                 # step names a segment that does not exist
    with pytest.raises(NativeError):
        ingest.read_gfa_many(items[:2] + [("chr2:0-1", str(bad), 1)], threads=3)


@pytest.mark.parametrize("text", [
    # dense numeric names (direct index), P and W lines, a revisit
    "S\t5\tACG\nS\t6\t*\tLN:i:7\nS\t8\tT\nP\ta#1#c\t5+,8-,5+\t*\nW\ts\t2\tc\t0\t9\t>6<8\n",
    # '7' and '007' are different segments: not canonical -> hash path
    "S\t7\tA\nS\t007\tCC\nS\t70\tGGG\nP\tp\t007+,70-\t*\nP\tq\t7+\t*\n",
    # sparse numbers (range far wider than the segment count) -> hash path
    "S\t1\tA\nS\t900000000\tCC\nP\tp\t900000000+,1+\t*\n",
    # ten-digit and non-numeric names
    "S\t1234567890\tA\nS\tutg1\tCC\nS\t3\tG\nW\ts\t1\tc\t*\t*\t>utg1>1234567890<3\n",
    # numeric file whose paths visit nothing / everything
    "S\t0\tA\nS\t1\tC\nS\t2\tG\nP\tempty\t*\t*\nP\tall\t0+,1+,2+\t*\n",
])
def test_segment_index_paths_agree_with_oracle(text):
    check(text)


@pytest.mark.parametrize("text", [
    "S\t1\tA\nS\t2\tC\nP\tp\t1+,02+\t*\n",          # '02' is not the name of any segment of this (numeric) file
    "S\t1\tA\nS\t2\tC\nP\tp\t1+,3+\t*\n",           # undefined number inside the range's neighbourhood
    "S\t1\tA\nS\t2\tC\nP\tp\t1+,x+\t*\n",           # non-numeric step in a numeric file
    "S\t10\tA\nS\t12\tC\nP\tp\t11+\t*\n",           # hole in the numbering
])
def test_undefined_steps_are_errors(text):
    with pytest.raises(NativeError):
        ingest.parse_gfa(text)
    with pytest.raises(KeyError):
        ogfa.parse(text)


# ---------------------------------------------------------------------------------- column compaction (ingest)
def _rand_window(rng, n, m, p_const=0.3, p_empty=0.1, p_zero=0.1, heavy=False):
    x = (rng.random((n, m)) < rng.random(m)[None, :]).astype(np.uint8)
    kind = rng.random(m)
    x[:, kind < p_const] = 1
    x[:, (kind >= p_const) & (kind < p_const + p_empty)] = 0
    hi = 100000 if heavy else 60
    nl = rng.integers(1, hi, size=m).astype(np.uint32)
    nl[rng.random(m) < 0.5] = 1
    nl[rng.random(m) < p_zero] = 0
    return x, nl


@pytest.mark.parametrize("n,m,heavy", [(1, 5, False), (2, 1, False), (7, 33, False), (40, 300, True), (130, 1009, False),
                                       (466, 1009, True), (5, 0, False), (0, 9, False)])
def test_compact_columns_matches_oracle_and_keeps_every_count(n, m, heavy):
    rng = np.random.default_rng(n * 1000 + m)
    x, nl = _rand_window(rng, n, m, heavy=heavy)
    win = ingest.GraphWindow([f"h{i}" for i in range(n)], similarity.pack_bits(x) if (m and n) else np.zeros((n, 4), np.uint32), nl)
    got = ingest.compact_window(win, pairs=False)
    x2, nl2 = similarity.compact_columns(x, nl)
    assert got.m == x2.shape[1] and got.x_bits.shape[1] % 4 == 0
    assert np.array_equal(got.node_len.astype(np.int64), nl2)
    if n:
        assert np.array_equal(similarity.unpack_bits(got.x_bits, got.m), x2)
        assert not similarity.unpack_bits(got.x_bits, got.x_bits.shape[1] * 32)[:, got.m:].any()  # padding bits are clear
    # the contract: intersections, path lengths, pi_ij and segregating nodes are those of the original window
    a, b = similarity.pairwise(x, nl), similarity.pairwise(x2, nl2)
    assert np.array_equal(a["I"], b["I"]) and np.array_equal(a["A"], b["A"]) and np.array_equal(a["pi"], b["pi"])
    assert similarity.segregating_nodes(x, nl) == similarity.segregating_nodes(x2, nl2)
    if n >= 4:
        rows = np.arange(0, n, 2)
        assert similarity.segregating_nodes(x, nl, rows) == similarity.segregating_nodes(x2, nl2, rows)
    if got.m:
        assert (np.diff(got.node_len[:got.m - 1].astype(np.int64)) >= 0).all()                      # ordered by length


def _bubble_window(rng, n, sites, extra, heavy=False, dup=True):
    """A window with the shapes the affine compaction acts on: backbone nodes, bi-allelic bubbles (complementary
    columns), columns in perfect linkage (identical), a tri-allelic site and unrelated columns."""
    cols, lens = [np.ones(n, np.uint8)], [int(rng.integers(1, 500))]
    for s_ in range(sites):
        alt = (rng.random(n) < rng.random()).astype(np.uint8)
        cols += [1 - alt, alt, np.ones(n, np.uint8)]
        lens += [int(rng.integers(1, 30)), int(rng.integers(1, 100000 if heavy and s_ % 5 == 0 else 60)), int(rng.integers(0, 200))]
        if dup and s_ % 4 == 0:                            # a second variant carried by exactly the same haplotypes
            cols += [1 - alt, alt]
            lens += [1, 1]
    a = rng.integers(0, 3, size=n)                         # three alleles: no two columns are complementary
    cols += [(a == 0).astype(np.uint8), (a == 1).astype(np.uint8), (a == 2).astype(np.uint8)]
    lens += [2, 3, 4]
    for _ in range(extra):
        cols.append((rng.random(n) < rng.random()).astype(np.uint8))
        lens.append(int(rng.integers(0, 40)))
    x = np.stack(cols, axis=1) if n else np.zeros((0, len(cols)), np.uint8)
    return x, np.array(lens, dtype=np.uint32)


@pytest.mark.parametrize("n,sites,extra,heavy,replicate", [(2, 3, 2, False, True), (3, 1, 0, False, True), (17, 20, 5, False, True),
                                                           (60, 90, 10, True, True), (60, 90, 10, True, False), (466, 280, 0, True, True),
                                                           (130, 40, 300, False, True), (1, 4, 3, False, True)])
def test_affine_compaction_is_exact(n, sites, extra, heavy, replicate):
    """impop_compact_fill with IMPOP_COMPACT_PAIRS (identical columns merged, the two branches of a bubble in one column,
    constants in C, optionally weights spread over copies): intersections, path lengths, unions, pi_ij and the
    segregating-node count of the affine form equal those of the original window, for every subset of rows."""
    rng = np.random.default_rng(n * 7919 + sites * 31 + extra)
    x, nl = _bubble_window(rng, n, sites, extra, heavy=heavy)
    win = ingest.GraphWindow([f"h{i}" for i in range(n)], similarity.pack_bits(x), nl)
    got = ingest.compact_window(win, replicate=replicate)
    x2 = similarity.unpack_bits(got.x_bits, got.m)
    assert not similarity.unpack_bits(got.x_bits, got.x_bits.shape[1] * 32)[:, got.m:].any()
    a, b = similarity.pairwise(x, nl), similarity.pairwise_affine(x2, got.node_len, got.row_adj, got.win_const)
    for key in ("I", "A", "U"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["pi"], b["pi"])
    assert similarity.segregating_nodes(x, nl) == similarity.segregating_nodes_affine(x2, got.node_len, got.col_mult)
    if n >= 4:
        rows = np.arange(0, n, 2)
        assert similarity.segregating_nodes(x, nl, rows) == similarity.segregating_nodes_affine(x2, got.node_len, got.col_mult, rows)
    assert got.site_runs == similarity.site_runs(x, nl)
    # fully reduced: no constant / empty / zero-length column is left, and (before weights are spread over copies) no two
    # columns are identical or complementary
    cnt = x2.astype(np.int64).sum(axis=0)
    assert ((cnt > 0) & (cnt < max(n, 1))).all() or n <= 1
    assert (got.node_len[:got.m] > 0).all()
    if n > 1:
        keys = {bytes(c) for c in x2.T[got.col_mult[:got.m] > 0]}
        assert len(keys) == int((got.col_mult[:got.m] > 0).sum())
        assert not any(bytes(1 - c) in keys for c in x2.T)
    if sites >= 3 and n >= 17 and extra <= sites:
        assert got.m < x.shape[1] // 2                     # bubbles and backbone gone: well under half the columns
    if replicate and heavy and n >= 17:
        # every weight the copies could absorb inside the chunk padding is a byte weight now
        assert (got.col_mult[:got.m] == 0).any()
    legacy = ingest.compact_window(win, pairs=False)
    assert legacy.row_adj is None and legacy.m >= got.m - int((got.col_mult[:got.m] == 0).sum())


def test_affine_compaction_leaves_incomplete_bubbles_alone():
    """A haplotype that visits neither branch (a gap in its assembly) or both breaks the complement: the columns stay."""
    x = np.array([[1, 0, 1], [0, 1, 1], [0, 0, 1], [1, 0, 0]], dtype=np.uint8)
    nl = np.array([5, 7, 11], dtype=np.uint32)
    got = ingest.compact_window(ingest.GraphWindow(list("abcd"), similarity.pack_bits(x), nl))
    assert got.m == 3 and got.win_const == 0 and not got.row_adj.any() and got.col_mult.tolist() == [1, 1, 1]
    x[2, 1] = 1                                             # now columns 0 and 1 are complementary
    got = ingest.compact_window(ingest.GraphWindow(list("abcd"), similarity.pack_bits(x), nl))
    assert got.m == 2 and got.win_const == 7 and sorted(got.col_mult.tolist()) == [1, 2]
    assert np.array_equal(similarity.pairwise(x, nl)["I"],
                          similarity.pairwise_affine(similarity.unpack_bits(got.x_bits, got.m), got.node_len, got.row_adj, got.win_const)["I"])


def test_compact_range_error():
    x = np.array([[1, 0], [0, 1]], dtype=np.uint8)
    nl = np.array([1 << 30, 1 << 30], dtype=np.uint32)
    with pytest.raises(NativeError):
        ingest.compact_window(ingest.GraphWindow(["a", "b"], similarity.pack_bits(x), nl))


def test_compact_uniform_batch_threads():
    ws = synth.make_windows(60, 20000, 5, seed=91)
    aff = ingest.compact_uniform(ws.x_bits, ws.node_len, threads=3)
    assert aff.row_adj.shape == (5, 60) and aff.win_const.shape == (5,) and aff.col_mult.shape == aff.node_len.shape
    for w in range(5):
        x2 = similarity.unpack_bits(aff.x[w], int(aff.m[w]))
        a = similarity.pairwise(ws.dense(w), ws.node_len[w])
        b = similarity.pairwise_affine(x2, aff.node_len[w, :aff.m[w]], aff.row_adj[w], aff.win_const[w])
        assert np.array_equal(a["I"], b["I"]) and np.array_equal(a["pi"], b["pi"])
        assert aff.m[w] <= (ws.m - 1) // 3 + 64           # one column per segregating site (+ copies of split weights)
    plain = ingest.compact_uniform(ws.x_bits, ws.node_len, threads=3, pairs=False)
    xo, lo, mo = plain.x, plain.node_len, plain.m
    assert plain.row_adj is None
    assert xo.shape[0] == 5 and xo.shape[1] == 60 and lo.shape[1] == xo.shape[2] * 32
    K = (ws.m - 1) // 3
    for w in range(5):
        x2, nl2 = similarity.compact_columns(ws.dense(w), ws.node_len[w])
        assert mo[w] == x2.shape[1] <= 2 * K + 1
        assert np.array_equal(similarity.unpack_bits(xo[w], int(mo[w])), x2)
        assert np.array_equal(lo[w, :mo[w]].astype(np.int64), nl2) and not lo[w, mo[w]:].any()


# ---------------------------------------------------------------------------------- flat window container
def test_flat_container_round_trip_and_labels(tmp_path):
    ws = synth.make_windows(30, 20000, 5, seed=5)
    wins = []
    for w in range(5):
        names = synth.haplotype_names(30, "chr2", 1000 * w, 1000 * w + 20000)
        if w == 3:
            names = names[::-1]                           # another row order in one window
        wins.append(ingest.GraphWindow(names, ws.x_bits[w], ws.node_len[w, :ws.m].copy(), None,
                                       f"CHM13#0#chr2:{1000 * w}-{1000 * w + 20000}", 20000))
    wins.append(ingest.GraphWindow(["plain_name", "other"], np.zeros((2, 4), np.uint32), np.array([3, 0, 7], np.uint32), None, None, 0))
    path = tmp_path / "w.impw"
    ingest.save_flat(path, wins)
    for mmap in (True, False):
        fb = ingest.load_flat(path, mmap=mmap)
        assert fb.windows == 6 and len(fb.uniq) == 32
        for w, g in enumerate(wins):
            h = fb.window(w)
            assert h.names == g.names and h.region == g.region and h.length == g.length
            assert np.array_equal(h.x_bits, g.x_bits) and np.array_equal(h.node_len, g.node_len)
        lab = fb.labels(pop_a=["S00000#1#", "S00001#"], pop_b=["S00002#2#"], subset=["S0000"])
        for w, g in enumerate(wins):
            pa = {s for s in g.names if s.startswith(("S00000#1#", "S00001#"))}
            pb = {s for s in g.names if s.startswith("S00002#2#")}
            sub = {s for s in g.names if s.startswith("S0000")}
            want = ingest.labels_from_names(g.names, pa, pb, sub, sub)
            assert np.array_equal(lab[int(fb.row_off[w]):int(fb.row_off[w]) + g.n], want), w
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"\0" * 128)
        ingest.load_flat(tmp_path / "bad")


# ---------------------------------------------------------------------------------- variant sites (bubble-like runs)
def test_site_runs_on_hand_built_graphs():
    """S as a bubble caller would count it (run_tajd.sh:126-148 counts `povu gfa2vcf` records): the
    ingest step's count against the restatement, on graphs whose sites are known by construction."""
    #            backbone | SNP bubble (ref, alt) | backbone | nested / adjacent alleles without backbone between | bb | absent node | bb
    x = np.array([[1, 1, 0, 1, 1, 0, 0, 1, 0, 1],
                  [1, 0, 1, 1, 0, 1, 0, 1, 0, 1],
                  [1, 1, 0, 1, 0, 0, 1, 1, 0, 1],
                  [1, 1, 0, 1, 1, 0, 0, 1, 0, 1]], dtype=np.uint8)
    nl = np.array([10, 1, 1, 20, 3, 4, 5, 30, 9, 2], dtype=np.uint32)
    assert similarity.site_runs(x, nl) == 2 and similarity.segregating_nodes(x, nl) == 5
    nl0 = nl.copy(); nl0[3] = 0                               # a zero-length backbone node does not separate: one site
    assert similarity.site_runs(x, nl0) == 1
    for xx, ll in ((x, nl), (x, nl0), (x[:1], nl), (np.ones((3, 4), np.uint8), np.ones(4, np.uint32)), (np.zeros((3, 4), np.uint8), np.ones(4, np.uint32))):
        win = ingest.GraphWindow([f"h{i}" for i in range(xx.shape[0])], similarity.pack_bits(xx), ll)
        assert ingest.compact_window(win).site_runs == similarity.site_runs(xx, ll)
    ws = synth.make_windows(40, 20000, 3, seed=12)
    K = (ws.m - 1) // 3
    for w in range(3):
        d = ws.dense(w)[:, :ws.m]
        poly = sum(1 for s in range(K) if 0 < d[:, 2 + 3 * s].sum() < 40)          # sites whose alt allele segregates
        assert similarity.site_runs(d, ws.node_len[w, :ws.m]) == poly
        assert 2 * poly == similarity.segregating_nodes(d, ws.node_len[w, :ws.m])


def test_containers_carry_the_affine_form(tmp_path):
    """A compacted window (affine form) survives both containers: row terms, window constant, column multiplicities and the
    variant-site count taken on the original node order come back as they went in; a plain window reads back as R = 0,
    C = 0, multiplicity 1."""
    ws = synth.make_windows(30, 20000, 3, seed=8)
    wins = []
    for w in range(3):
        g = ingest.GraphWindow(synth.haplotype_names(30, "chr2", 1000 * w, 1000 * w + 20000), ws.x_bits[w], ws.node_len[w, :ws.m].copy(),
                               None, f"chr2:{1000 * w}-{1000 * w + 20000}", 20000)
        wins.append(ingest.compact_window(g) if w != 1 else g)               # window 1 stays plain
    assert wins[0].row_adj is not None and wins[0].win_const > 0 and wins[1].row_adj is None
    ingest.save_flat(tmp_path / "w.impw", wins)
    ingest.save_batch(tmp_path / "w.npz", wins)
    fb = ingest.load_flat(tmp_path / "w.impw")
    for back in ([fb.window(w) for w in range(3)], ingest.load_batch(tmp_path / "w.npz")):
        for g, h in zip(wins, back):
            assert np.array_equal(h.x_bits, g.x_bits) and np.array_equal(h.node_len, g.node_len) and h.names == g.names
            assert h.site_runs == g.site_runs
            if g.row_adj is None:
                assert not np.asarray(h.row_adj).any() and h.win_const == 0 and (np.asarray(h.col_mult) == 1).all()
            else:
                assert np.array_equal(h.row_adj, g.row_adj) and h.win_const == g.win_const and np.array_equal(h.col_mult, g.col_mult)
    he = ingest.heavy_entries(np.array([[0, 254, 255, 510, 65024, 65025, 65026 + 255]], dtype=np.uint32))
    assert he.tolist() == [0 + 0 + 1 + 1 + 1 + 1 + 2]


def test_compaction_plan_cache_is_guarded_by_content():
    """impop_compact_fill reuses the plans of the preceding impop_compact_scan only for the same CONTENT: a batch changed in
    place between the two calls (same addresses, same sizes) is planned again."""
    from impop_b200._native import lib, COMPACT_PAIRS, COMPACT_REPLICATE
    L = lib()
    rng = np.random.default_rng(3)
    xa, nla = _bubble_window(rng, 40, 12, 4)
    xb, nlb = _bubble_window(rng, 40, 12, 4)
    assert xa.shape == xb.shape
    bits = similarity.pack_bits(xa).copy()
    nl = nla.copy()
    n, m, pitch = np.array([40], np.int32), np.array([xa.shape[1]], np.int32), np.array([bits.shape[1]], np.int32)
    zero = np.zeros(1, np.int64)
    flags = COMPACT_PAIRS | COMPACT_REPLICATE
    m_out, runs = np.zeros(1, np.int32), np.zeros(1, np.int64)
    assert L.impop_compact_scan(1, n.ctypes.data, m.ctypes.data, pitch.ctypes.data, zero.ctypes.data, zero.ctypes.data, bits.ctypes.data,
                                nl.ctypes.data, 1, flags, m_out.ctypes.data, runs.ctypes.data) == 0
    bits[:] = similarity.pack_bits(xb)                      # another window at the same addresses
    nl[:] = nlb
    want = ingest.compact_window(ingest.GraphWindow([f"h{i}" for i in range(40)], similarity.pack_bits(xb), nlb))
    op = np.array([max(4, (max(int(m_out[0]), want.m) + 127) // 128 * 4)], np.int32)
    x_out, len_out = np.zeros(40 * int(op[0]), np.uint32), np.zeros(32 * int(op[0]), np.uint32)
    ra, wc, cmult = np.zeros(40, np.int32), np.zeros(1, np.int64), np.zeros(32 * int(op[0]), np.uint8)
    assert L.impop_compact_fill(1, n.ctypes.data, m.ctypes.data, pitch.ctypes.data, zero.ctypes.data, zero.ctypes.data, bits.ctypes.data,
                                nl.ctypes.data, 1, flags, op.ctypes.data, zero.ctypes.data, zero.ctypes.data, x_out.ctypes.data,
                                len_out.ctypes.data, zero.ctypes.data, ra.ctypes.data, wc.ctypes.data, cmult.ctypes.data) == 0
    assert np.array_equal(len_out[:want.m], want.node_len) and not len_out[want.m:].any()
    assert int(wc[0]) == want.win_const and np.array_equal(ra, want.row_adj)
    assert np.array_equal(x_out.reshape(40, -1)[:, :want.x_bits.shape[1]], want.x_bits)


def test_compact_windows_equals_one_by_one():
    """ingest.compact_windows (one native call over a ragged list, all host threads) gives what compact_window gives."""
    rng = np.random.default_rng(17)
    wins = []
    for n, sites, extra in ((5, 3, 2), (40, 30, 6), (1, 2, 0), (130, 60, 40), (0, 1, 1)):
        x, nl = _bubble_window(rng, n, sites, extra, heavy=n > 30)
        wins.append(ingest.GraphWindow([f"h{i}" for i in range(n)], similarity.pack_bits(x) if n else np.zeros((0, 4), np.uint32), nl, None, f"r{n}", 1000 + n))
    for pairs in (True, False):
        many = ingest.compact_windows(wins, threads=3, pairs=pairs)
        for w, g in zip(wins, many):
            one = ingest.compact_window(w, pairs=pairs)
            assert g.m == one.m and g.names == one.names and g.region == one.region and g.length == one.length and g.site_runs == one.site_runs
            assert np.array_equal(g.x_bits, one.x_bits) and np.array_equal(g.node_len, one.node_len)
            if pairs:
                assert np.array_equal(g.row_adj, one.row_adj) and g.win_const == one.win_const and np.array_equal(g.col_mult, one.col_mult)
            else:
                assert g.row_adj is None
    assert ingest.compact_windows([]) == []


def test_affine_compaction_with_hundreds_of_identical_columns():
    """More identical columns than one multiplicity byte can count: they are merged in groups of at most 127 (a pair of
    complementary groups: 254), and every count stays exact."""
    rng = np.random.default_rng(5)
    n = 20
    a = (rng.random(n) < 0.4).astype(np.uint8)
    a[0], a[1] = 0, 1
    cols = [a] * 300 + [1 - a] * 290 + [(rng.random(n) < 0.5).astype(np.uint8)]
    x = np.stack(cols, axis=1)
    nl = rng.integers(1, 9, size=x.shape[1]).astype(np.uint32)
    got = ingest.compact_window(ingest.GraphWindow([f"h{i}" for i in range(n)], similarity.pack_bits(x), nl))
    x2 = similarity.unpack_bits(got.x_bits, got.m)
    assert got.col_mult.max() <= 254 and int(got.col_mult.sum()) == x.shape[1]        # every node of positive length is counted once
    a0, b0 = similarity.pairwise(x, nl), similarity.pairwise_affine(x2, got.node_len, got.row_adj, got.win_const)
    assert np.array_equal(a0["I"], b0["I"]) and np.array_equal(a0["A"], b0["A"]) and np.array_equal(a0["pi"], b0["pi"])
    assert similarity.segregating_nodes(x, nl) == similarity.segregating_nodes_affine(x2, got.node_len, got.col_mult)
    assert int((got.col_mult > 0).sum()) <= 6              # 3 + 3 groups, one pair merged, + the unrelated column (the rest: copies of split weights)


def test_reader_reports_revisits_and_auto_counts():
    """impop_gfa_fill tells whether some path visits a node more than once; want_counts="auto" then reads the visit counts
    only for such windows (P and W lines, named and numbered segments, a revisit far from the first visit)."""
    plain = "S\t1\tAC\nS\t2\tG\nS\t3\tTT\nP\ta\t1+,2+,3+\t*\nP\tb\t1+,3+\t*\n"
    loop = "S\t1\tAC\nS\t2\tG\nS\t3\tTT\nP\ta\t1+,2+,3+,1-\t*\nP\tb\t1+,3+\t*\n"
    walk = "S\ts1\tAC\nS\ts2\tG\nW\tx\t1\tc\t0\t5\t>s1>s2<s1\n"
    big = "".join(f"S\t{k}\tA\n" for k in range(1, 200)) + "P\tp\t" + ",".join(f"{k}+" for k in list(range(1, 200)) + [7]) + "\t*\n"
    for text, want in ((plain, False), (loop, True), (walk, True), (big, True)):
        g = ingest.parse_gfa(text)
        assert g.revisits is want and g.counts is None
        a = ingest.parse_gfa(text, want_counts="auto")
        assert a.revisits is want and (a.counts is not None) is want
        if want:
            full = ingest.parse_gfa(text, want_counts=True)
            assert np.array_equal(a.counts, full.counts) and int(a.counts.max()) == 2
        assert np.array_equal(a.x_bits, g.x_bits)
