"""Windows as column ranges of one chromosome-scale matrix against per-window batches and the
CPU oracle: aligned and unaligned range starts, overlapping (sliding) windows, empty windows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from impop_b200 import synth  # noqa: E402
from impop_b200.chromosome import ChromosomeMatrix, concatenate_windows  # noqa: E402
from oracle import clib, similarity  # noqa: E402
from oracle.compare import row_mismatches  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from impop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _labels(n):
    lab = np.full(n, 1 | 8, dtype=np.uint8)
    lab[: n // 3] |= 2
    lab[n // 3: 2 * n // 3] |= 4
    return lab


@pytest.mark.parametrize("algo", [0, 1])
def test_windows_of_a_concatenated_chromosome_match_their_own_batches(ctx, algo):
    from impop_b200.engine import WindowBatch
    ws = synth.make_windows(70, 20000, 6, seed=0xB200 + 21, n_sites_override=50)     # m = 151: unaligned range starts
    lab = _labels(ws.n)
    x_bits, node_len, pos = concatenate_windows(ws)
    chrom = ChromosomeMatrix(ctx, x_bits, node_len, pos)
    starts = np.arange(ws.windows) * ws.length
    batch = chrom.windows(starts, starts + ws.length, lab)
    got_s, got_c = batch.stats(algo)
    ref = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, lab, ws.length)
    want_s, want_c = ref.stats(algo)
    ctx.check()
    assert torch.equal(got_c, want_c)
    assert torch.equal(got_s.nan_to_num(7.0), want_s.nan_to_num(7.0))      # same pairs, same summation order: same bits
    batch.close(); ref.close()


def test_sliding_windows_against_the_oracle(ctx):
    ws = synth.make_windows(48, 10000, 8, seed=0xB200 + 22, n_sites_override=30)
    lab = _labels(ws.n)
    x_bits, node_len, pos = concatenate_windows(ws)
    chrom = ChromosomeMatrix(ctx, x_bits, node_len, pos)
    starts, ends, batch = chrom.sliding(25000, 7000, lab, begin=0, end=ws.windows * ws.length)
    assert len(starts) == 8
    st, ct = batch.stats(0)
    ctx.check()
    st, ct = st.cpu().numpy(), ct.cpu().numpy()
    dense = similarity.unpack_bits(x_bits, len(node_len))
    for w, (s, e) in enumerate(zip(starts, ends)):
        k0, k1 = np.searchsorted(pos, s, "left"), np.searchsorted(pos, e, "left")
        sub = np.ascontiguousarray(dense[:, k0:k1])
        m_pad = ((k1 - k0 + 127) // 128) * 128
        nl = np.zeros(m_pad, dtype=np.uint32)
        nl[: k1 - k0] = node_len[k0:k1]
        want_s, want_c = clib.window_stats(similarity.pack_bits(sub, m_pad // 32), m_pad, nl, lab, int(e - s))
        assert (ct[w] == want_c).all(), (w, ct[w], want_c)
        assert not row_mismatches(st[w], want_s, 1e-12), w
    batch.close()


def test_empty_and_edge_windows(ctx):
    ws = synth.make_windows(10, 5000, 2, seed=3, n_sites_override=5)
    x_bits, node_len, pos = concatenate_windows(ws)
    chrom = ChromosomeMatrix(ctx, x_bits, node_len, pos)
    lab = np.full(10, 9, dtype=np.uint8)
    batch = chrom.windows([0, 20000, 4990], [5000, 30000, 5010], lab)       # whole first window, beyond the end, a straddler
    st, ct = batch.stats(0)
    ctx.check()
    st = st.cpu().numpy()
    assert st[1][0] == 0.0 or np.isnan(st[1][0]) or st[1][0] == 1.0         # no nodes: every path empty (J := 0 -> pi_ij = 1)
    dense = similarity.unpack_bits(x_bits, len(node_len))
    k0, k1 = np.searchsorted(pos, 4990, "left"), np.searchsorted(pos, 5010, "left")
    sub = np.ascontiguousarray(dense[:, k0:k1])
    nl = np.zeros(128, dtype=np.uint32); nl[: k1 - k0] = node_len[k0:k1]
    want_s, want_c = clib.window_stats(similarity.pack_bits(sub, 4), 128, nl, lab, 20)
    assert not row_mismatches(st[2], want_s, 1e-12)
    batch.close()
