"""Matrix mode from window graphs: GFA text -> libimpop_b200 reader -> fused GPU pass -> the
wrappers' TSVs, against the CPU oracle run on the generator's own matrices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from impop_b200 import ingest, synth, windows  # noqa: E402
from oracle import clib  # noqa: E402
from oracle.compare import row_mismatches  # noqa: E402


def test_gfa_windows_through_the_driver(tmp_path):
    n, L, W = 60, 50000, 4
    ws = synth.make_windows(n, L, W, seed=0xB200 + 7, n_sites_override=60)
    idx = np.arange(n)
    names = synth.haplotype_names(n, "chr2", 0, L)
    asm = synth.assembly_names(idx)
    pop_a, pop_b = [asm[i] for i in range(0, 20)], [asm[i] for i in range(20, 45)]
    (tmp_path / "a.txt").write_text("\n".join(pop_a) + "\n")
    (tmp_path / "b.txt").write_text("# population B\n" + "\n".join(pop_b) + "\n")
    listing = []
    for w in range(W):
        region = windows.region_name("chr2", w * L, (w + 1) * L)
        path = tmp_path / f"w{w}.gfa"
        with open(path, "w") as fh:
            ingest.write_gfa(fh, names, ws.dense(w)[:, :ws.m], ws.node_len[w, :ws.m], walks=bool(w & 1))
        listing.append(f"{region}\t{path}")
    (tmp_path / "windows.tsv").write_text("\n".join(listing) + "\n")
    rc = windows.main(["--gfa-list", str(tmp_path / "windows.tsv"), "-a", str(tmp_path / "a.txt"), "-b", str(tmp_path / "b.txt"),
                       "--fst-out", str(tmp_path / "fst.tsv"), "--tajd-out", str(tmp_path / "tajd.tsv"),
                       "--pi-out", str(tmp_path / "pi.tsv"), "--save-batch", str(tmp_path / "batch.npz"),
                       "--pooled-fst-out", str(tmp_path / "pooled.tsv")])
    assert rc == 0
    # row a-10 (run_fst_impg.sh:158-220): pica2's printed per-site pi of A, B, A + B and the pooled estimator on the text
    from oracle import popstats
    pooled = (tmp_path / "pooled.tsv").read_text().splitlines()
    assert pooled[0].split("\t") == ["REGION", "LENGTH", "THRESHOLD", "R_VALUE", "PI_A", "PI_B", "PI_C", "PI_AB_AVG", "FST"]
    for w in range(W):
        txt = []
        for rows in (range(0, 20), range(20, 45), range(0, 45)):
            sub_lab = np.zeros(n, dtype=np.uint8)
            sub_lab[list(rows)] = 1 | 8
            txt.append(f"{clib.window_stats(ws.x_bits[w], ws.m_pad, ws.node_len[w], sub_lab, L)[0][1]:.8f}")
        avg, fst_txt = popstats.pooled_fst_text(*txt)
        assert pooled[1 + w].split("\t")[4:] == txt + [avg, fst_txt], (w, pooled[1 + w])
    # the rows the reference-side restatement gives on the generator's matrices
    lab = np.full(n, 1 | 8, dtype=np.uint8)
    lab[0:20] |= 2            # asm[i] is the assembly-name spelling of names[i]
    lab[20:45] |= 4
    fst = (tmp_path / "fst.tsv").read_text().splitlines()
    taj = (tmp_path / "tajd.tsv").read_text().splitlines()
    pi = (tmp_path / "pi.tsv").read_text().splitlines()
    assert fst[0].split("\t") == windows.HEADERS["fst"] and taj[0].split("\t") == windows.HEADERS["tajd"]
    for w in range(W):
        want_s, want_c = clib.window_stats(ws.x_bits[w], ws.m_pad, ws.node_len[w], lab, L)
        f = fst[1 + w].split("\t")
        assert f[0] == windows.region_name("chr2", w * L, (w + 1) * L) and f[1] == str(L)
        got = np.array([float(v) for v in f[2:]])
        assert np.allclose(got, [want_s[k] for k in (7, 2, 3, 4, 5, 6)], rtol=0, atol=6e-9)      # %.8f text
        t = taj[1 + w].split("\t")
        assert t[2] == str(n) and t[3] == str(int(want_c[7]))
        assert pi[1 + w].split("\t")[-1] == f"{want_s[1]:.8f} (sequence length: {L})"
    # the binary container reproduces the same statistics bit for bit
    from impop_b200.engine import Context
    ctx = Context(0)
    graphs = ingest.load_batch(tmp_path / "batch.npz")
    batch = windows.batch_from_graphs(ctx, graphs, set(pop_a), set(pop_b))
    st, ct = batch.stats()
    ctx.check()
    st, ct = st.cpu().numpy(), ct.cpu().numpy()
    for w in range(W):
        want_s, want_c = clib.window_stats(ws.x_bits[w], ws.m_pad, ws.node_len[w], lab, L)
        assert not row_mismatches(st[w], want_s, 1e-12)
        assert (ct[w] == want_c).all()
    batch.close()
    ctx.close()


def test_multiset_coverage_window_on_the_gpu():
    """f-4: a window whose paths revisit nodes, expanded to copy nodes, through the fused pass; expected values from the
    min-count definition directly (oracle.similarity.identity_from_counts on I = sum len min(c_i, c_j), A = sum len c)."""
    from impop_b200.engine import Context, WindowBatch
    from oracle import similarity
    rng = np.random.default_rng(10)
    n, m = 40, 90
    counts = rng.integers(0, 3, size=(n, m))
    counts[:, ::7] = 1
    node_len = rng.integers(1, 500, size=m)
    lines = ["H\tVN:Z:1.0"] + [f"S\t{k + 1}\t*\tLN:i:{node_len[k]}" for k in range(m)]
    for i in range(n):
        steps = [f"{k + 1}+" for k in range(m) for _ in range(counts[i, k])]
        lines.append(f"P\tS{i:03d}#1#ctg:0-1000\t" + ",".join(steps) + "\t*")
    win = ingest.multiset_expand(ingest.parse_gfa("\n".join(lines) + "\n", want_counts=True))
    inter = np.einsum("k,ijk->ij", node_len.astype(np.int64), np.minimum(counts[:, None, :], counts[None, :, :]))
    a = counts @ node_len
    _, _, _, pi = similarity.identity_from_counts(inter, a)
    ctx = Context(0)
    batch = WindowBatch.from_windows(ctx, [(win.x_bits, win.node_len, np.full(n, 9, dtype=np.uint8), 1000)])
    I, A, P = batch.pairwise(0, 0)
    ctx.check()
    assert np.array_equal(I.cpu().numpy(), inter) and np.array_equal(A.cpu().numpy(), a)
    got = P.cpu().numpy()
    iu = np.triu_indices(n, 1)
    assert np.array_equal(got[iu], pi[iu])                     # bit-exact pi_ij
    st, ct = batch.stats(0)
    ctx.check()
    want_pi = n / (n - 1) * 2 * float(np.sum(pi[iu])) / (n * n)
    assert abs(float(st[0, 0]) - want_pi) <= 1e-12 * want_pi
    batch.close(); ctx.close()


def _write_listing(tmp_path, names, ws, W, L):
    listing = []
    for w in range(W):
        path = tmp_path / f"w{w}.gfa"
        with open(path, "w") as fh:
            ingest.write_gfa(fh, names, ws.dense(w)[:, :ws.m], ws.node_len[w, :ws.m])
        listing.append(f"{windows.region_name('chr2', w * L, (w + 1) * L)}\t{path}")
    (tmp_path / "windows.tsv").write_text("\n".join(listing) + "\n")


def test_flat_container_gives_the_same_tables(tmp_path):
    """--save-batch x.impw then --batch x.impw (memory-mapped, one upload, labels per unique haplotype): byte-identical TSVs;
    the ingest-time compaction (default) and --no-compact agree too."""
    n, L, W = 50, 20000, 6
    ws = synth.make_windows(n, L, W, seed=0xB200 + 9, n_sites_override=40)
    names = synth.haplotype_names(n, "chr2", 0, L)
    asm = synth.assembly_names(np.arange(n))
    (tmp_path / "a.txt").write_text("\n".join(asm[:15]) + "\n")
    (tmp_path / "b.txt").write_text("\n".join(asm[10:30]) + "\n")          # overlaps A: dropped from both (h-fst.py:181-185)
    (tmp_path / "s.txt").write_text("\n".join(asm[:40]) + "\n")
    _write_listing(tmp_path, names, ws, W, L)
    common = ["-a", str(tmp_path / "a.txt"), "-b", str(tmp_path / "b.txt"), "-s", str(tmp_path / "s.txt")]

    def run(tag, *src):
        outs = [str(tmp_path / f"{tag}.{k}.tsv") for k in ("pi", "fst", "tajd")]
        assert windows.main([*src, *common, "--pi-out", outs[0], "--fst-out", outs[1], "--tajd-out", outs[2]]) == 0
        return [open(o).read() for o in outs]
    first = run("gfa", "--gfa-list", str(tmp_path / "windows.tsv"), "--save-batch", str(tmp_path / "b.impw"))
    assert run("flat", "--batch", str(tmp_path / "b.impw")) == first
    assert run("raw", "--gfa-list", str(tmp_path / "windows.tsv"), "--no-compact") == first
    # the D column follows from the PI column of the same row (run_tajd.sh:180 hands tj_d.py the printed pi)
    from oracle import popstats
    for line in first[2].splitlines()[1:]:
        f = line.split("\t")
        d, _ = popstats.tajimas_d(int(f[2]), float(f[3]), float(f[4]))
        assert f[5] == ("NA" if d != d else repr(d)), line


def test_revisited_nodes_are_multiset_by_default(tmp_path):
    """a path that visits a node twice contributes min(count_a, count_b) * len by default; --presence-only opts out."""
    gfa = ("H\tVN:Z:1.0\nS\t1\t*\tLN:i:100\nS\t2\t*\tLN:i:40\nS\t3\t*\tLN:i:7\n"
           "P\tA#1#c:0-200\t1+,2+,2+,3+\t*\nP\tB#1#c:0-200\t1+,2+,3+\t*\nP\tC#1#c:0-200\t1+,2+,2+,2+\t*\n")
    (tmp_path / "w.gfa").write_text(gfa)
    (tmp_path / "l.tsv").write_text(f"CHM13#0#chr2:0-200\t{tmp_path / 'w.gfa'}\n")

    def pi_of(*extra):
        out = tmp_path / ("pi" + "".join(extra) + ".tsv")
        assert windows.main(["--gfa-list", str(tmp_path / "l.tsv"), "--pi-out", str(out), *extra]) == 0
        return float(out.read_text().splitlines()[1].split("\t")[-1].split()[0]) * 200        # per-site text -> pi

    def pi_from(I, A):
        from oracle import similarity
        _, _, _, pi = similarity.identity_from_counts(np.array(I), np.array(A))
        return 3 / 2 * 2 * (pi[0, 1] + pi[0, 2] + pi[1, 2]) / 9
    multi = pi_from([[187, 147, 180], [147, 147, 140], [180, 140, 220]], [187, 147, 220])
    pres = pi_from([[147, 147, 140], [147, 147, 140], [140, 140, 140]], [147, 147, 140])
    assert abs(pi_of() - multi) < 2e-6 and abs(pi_of("--presence-only") - pres) < 2e-6 and abs(multi - pres) > 1e-3


def test_disjoint_pairs_convention(tmp_path):
    """The fused path counts a pair of paths that share no node with pi_ij = 1; --disjoint-absent treats it as a row the
    similarity tool did not print (skipped and not counted, pica2.py:132-134, h-fst.py:147-153)."""
    from oracle import popstats, similarity
    rng = np.random.default_rng(4)
    n, m = 12, 40
    x = np.zeros((n, m), dtype=np.uint8)
    x[:7, :25] = rng.random((7, 25)) < 0.7
    x[7:, 25:] = rng.random((5, 15)) < 0.7
    x[:7, 0] = 1; x[7:, 25] = 1
    nl = rng.integers(1, 30, size=m)
    names = [f"H{i:02d}#1#c:0-500" for i in range(n)]
    with open(tmp_path / "w.gfa", "w") as fh:
        ingest.write_gfa(fh, names, x, nl)
    (tmp_path / "l.tsv").write_text(f"CHM13#0#chr2:0-500\t{tmp_path / 'w.gfa'}\n")
    res = similarity.pairwise(x, nl)
    ident = res["identity"].copy()
    counted = 2 * float(np.sum(res["pi"][np.triu_indices(n, 1)])) / (n * n) * n / (n - 1)
    ident[res["I"] == 0] = np.nan
    np.fill_diagonal(ident, np.nan)
    absent, _ = popstats.pica2_pi(ident, names, 1.0, None)
    outs = {}
    for tag, extra in (("counted", []), ("absent", ["--disjoint-absent"])):
        out = tmp_path / f"{tag}.tsv"
        assert windows.main(["--gfa-list", str(tmp_path / "l.tsv"), "--pi-out", str(out), *extra]) == 0
        outs[tag] = float(out.read_text().splitlines()[1].split("\t")[-1].split()[0]) * 500
    assert abs(outs["counted"] - counted) < 5e-6 and abs(outs["absent"] - absent) < 5e-6 and counted - absent > 0.05
