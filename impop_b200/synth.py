"""Synthetic "HPRC-shaped" window graphs (SURVEY.md section 8 d).

Host-side data generation only (numpy, seeded); nothing here is on the timed path.
A window is a presence matrix x[n haplotypes, m nodes] plus node lengths; nodes are
backbone_0, then (ref_s, alt_s, backbone_s) for each of K variant sites, padded with
zero-length all-absent nodes to a multiple of 128 columns.

Panel sizes follow doc/where_hprc_data.md:5-10 of the reference (2 x individuals,
trimmed to the 466 haplotypes BASELINE.json names).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

PANEL_466 = (("AFR", 140), ("AMR", 88), ("EAS", 100), ("EUR", 60), ("SAS", 72), ("UNK", 6))
THETA = 1e-3
FST_BN = 0.12


def harmonic(n: int) -> float:
    return float(sum(1.0 / i for i in range(1, n)))


def panel(n_total: int = 466, spec=PANEL_466):
    """Population label (index into spec) per haplotype, in panel order, scaled to n_total."""
    base = np.array([c for _, c in spec], dtype=np.float64)
    if n_total == int(base.sum()):
        counts = base.astype(int)
    else:
        counts = np.floor(base * n_total / base.sum()).astype(int)
        counts[0] += n_total - counts.sum()
    pops = np.repeat(np.arange(len(spec)), counts)
    return pops, [name for name, _ in spec]


def haplotype_names(n: int, chrom: str = "chr2", start: int = 0, end: int = 0):
    """PanSN-style ids `S00000#1#ctg0:start-end` (sample k = i // 2, hap 1/2)."""
    return [f"S{i // 2:05d}#{i % 2 + 1}#ctg{i // 2}:{start}-{end}" for i in range(n)]


def assembly_names(idx):
    """Population-list spelling of haplotype indices: `S00000_hap1_hprc_r2_v1.0.1` (exercises h-fst.py:18-61)."""
    return [f"S{i // 2:05d}_hap{i % 2 + 1}_hprc_r2_v1.0.1" for i in idx]


@dataclass
class WindowSet:
    """A batch of W same-shape windows: x_bits[W, n, pitch_words] u32, node_len[W, m_pad] u32."""
    x_bits: np.ndarray
    node_len: np.ndarray
    n: int
    m: int                 # real node count (3K+1)
    m_pad: int
    length: int            # BED window length L
    pops: np.ndarray       # (n,) population index
    pop_names: list = field(default_factory=list)

    @property
    def windows(self) -> int:
        return self.x_bits.shape[0]

    def dense(self, w: int) -> np.ndarray:
        bits = np.unpackbits(self.x_bits[w].view(np.uint8).reshape(self.n, -1), axis=1, bitorder="little")
        return bits[:, : self.m_pad]


def n_sites(n: int, length: int) -> int:
    return max(1, int(round(THETA * length * harmonic(n))))


def make_windows(n: int, length: int, windows: int, seed: int, n_sites_override: int | None = None,
                 chunk: int = 128, pops: np.ndarray | None = None, pop_names=None,
                 max_sv_len: int = 10000) -> WindowSet:
    """Generate `windows` windows of n haplotypes over a BED length `length`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if pops is None:
        pops, pop_names = panel(n)
    npop = int(pops.max()) + 1
    K = n_sites_override if n_sites_override is not None else n_sites(n, length)
    m = 3 * K + 1
    m_pad = ((m + 127) // 128) * 128
    pitch = m_pad // 32
    x_bits = np.empty((windows, n, pitch), dtype=np.uint32)
    node_len = np.zeros((windows, m_pad), dtype=np.uint32)
    sfs = 1.0 / np.arange(1, n)
    sfs /= sfs.sum()
    a_bn = (1.0 - FST_BN) / FST_BN
    ref_col = 1 + 3 * np.arange(K)
    for w0 in range(0, windows, chunk):
        wc = min(chunk, windows - w0)
        # --- site classes and node lengths
        cls = rng.random((wc, K))
        ref_len = np.ones((wc, K), dtype=np.int64)
        alt_len = np.ones((wc, K), dtype=np.int64)
        indel = (cls >= 0.90) & (cls < 0.98)
        sv = cls >= 0.98
        ind_len = rng.integers(1, 51, size=(wc, K))
        ins = rng.random((wc, K)) < 0.5
        alt_len = np.where(indel & ins, ind_len, alt_len)
        ref_len = np.where(indel & ~ins, ind_len, ref_len)
        sv_len = np.exp(rng.uniform(np.log(50.0), np.log(float(max_sv_len)), size=(wc, K))).astype(np.int64)
        alt_len = np.where(sv, sv_len, alt_len)
        # backbone lengths: multinomial so that sum(backbone + ref) == L
        spare = length - ref_len.sum(axis=1)
        spare = np.maximum(spare, 0)
        bb = np.stack([rng.multinomial(int(spare[i]), np.full(K + 1, 1.0 / (K + 1))) for i in range(wc)])
        nl = node_len[w0:w0 + wc]
        nl[:, 0] = bb[:, 0]
        nl[:, ref_col] = ref_len
        nl[:, ref_col + 1] = alt_len
        nl[:, ref_col + 2] = bb[:, 1:]
        # --- genotypes
        k_anc = rng.choice(np.arange(1, n), size=(wc, K), p=sfs)
        p_anc = k_anc / n
        p_pop = rng.beta(p_anc[:, None, :] * a_bn, (1.0 - p_anc[:, None, :]) * a_bn, size=(wc, npop, K))
        u = rng.random((wc, n, K), dtype=np.float32)
        alt = u < p_pop[:, pops, :].astype(np.float32)
        dense = np.zeros((wc, n, m_pad), dtype=np.uint8)
        dense[:, :, 0] = 1
        dense[:, :, ref_col] = ~alt
        dense[:, :, ref_col + 1] = alt
        dense[:, :, ref_col + 2] = 1
        packed = np.packbits(dense, axis=2, bitorder="little")
        x_bits[w0:w0 + wc] = packed.view("<u4").reshape(wc, n, pitch)
    return WindowSet(x_bits=x_bits, node_len=node_len, n=n, m=m, m_pad=m_pad, length=length,
                     pops=pops, pop_names=list(pop_names) if pop_names else [])


def make_site_matrix(sites: int, n: int, seed: int, pops: np.ndarray | None = None, chunk: int = 1 << 18):
    """BASELINE config 4: site-major bit matrix (sites x ceil(n/64) u64) + per-population masks."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if pops is None:
        pops, _ = panel(n)
    npop = int(pops.max()) + 1
    words = (n + 63) // 64
    out = np.empty((sites, words), dtype=np.uint64)
    sfs = 1.0 / np.arange(1, n)
    sfs /= sfs.sum()
    a_bn = (1.0 - FST_BN) / FST_BN
    for s0 in range(0, sites, chunk):
        sc = min(chunk, sites - s0)
        p_anc = rng.choice(np.arange(1, n), size=sc, p=sfs) / n
        p_pop = rng.beta(p_anc[:, None] * a_bn, (1.0 - p_anc[:, None]) * a_bn, size=(sc, npop))
        u = rng.random((sc, n), dtype=np.float32)
        alt = (u < p_pop[:, pops].astype(np.float32))
        dense = np.zeros((sc, words * 64), dtype=np.uint8)
        dense[:, :n] = alt
        out[s0:s0 + sc] = np.packbits(dense, axis=1, bitorder="little").view("<u8").reshape(sc, words)
    masks = np.zeros((npop, words), dtype=np.uint64)
    for p in range(npop):
        d = np.zeros(words * 64, dtype=np.uint8)
        d[:n] = pops == p
        masks[p] = np.packbits(d, bitorder="little").view("<u8")
    return out, masks


def make_windows_device(ctx, n: int, length: int, windows: int, seed: int, n_sites_override: int | None = None,
                        chunk: int = 512, pops: np.ndarray | None = None, max_sv_len: int = 10000):
    """The same window model generated ON the device with torch's RNG (bench-scale batches: the numpy
    generator needs ~16 ms per window).  Returns (x_bits int32 [W, n, pitch], node_len int32 [W, m_pad],
    pops numpy, m, m_pad) with the tensors on ctx.torch_device; rows are bit-packed by impop_pack_bits.
    Backbone lengths are a uniform random composition of the spare length (sum(backbone + ref) == L)."""
    import torch
    dev = ctx.torch_device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    if pops is None:
        pops, _ = panel(n)
    pops_t = torch.from_numpy(np.asarray(pops)).to(dev)
    npop = int(pops.max()) + 1
    K = n_sites_override if n_sites_override is not None else n_sites(n, length)
    m = 3 * K + 1
    m_pad = ((m + 127) // 128) * 128
    pitch = m_pad // 32
    x_bits = torch.empty((windows, n, pitch), dtype=torch.int32, device=dev)
    node_len = torch.zeros((windows, m_pad), dtype=torch.int32, device=dev)
    sfs = 1.0 / torch.arange(1, n, device=dev, dtype=torch.float64)
    sfs = sfs / sfs.sum()
    a_bn = (1.0 - FST_BN) / FST_BN
    ref_col = 1 + 3 * torch.arange(K, device=dev)
    for w0 in range(0, windows, chunk):
        wc = min(chunk, windows - w0)
        cls = torch.rand((wc, K), generator=gen, device=dev)
        indel = (cls >= 0.90) & (cls < 0.98)
        sv = cls >= 0.98
        ind_len = torch.randint(1, 51, (wc, K), generator=gen, device=dev)
        ins = torch.rand((wc, K), generator=gen, device=dev) < 0.5
        one = torch.ones((wc, K), dtype=torch.int64, device=dev)
        alt_len = torch.where(indel & ins, ind_len, one)
        ref_len = torch.where(indel & ~ins, ind_len, one)
        lo, hi = float(np.log(50.0)), float(np.log(float(max_sv_len)))
        sv_len = torch.exp(lo + (hi - lo) * torch.rand((wc, K), generator=gen, device=dev, dtype=torch.float64)).to(torch.int64)
        alt_len = torch.where(sv, sv_len, alt_len)
        spare = (length - ref_len.sum(dim=1)).clamp_min(0)
        cuts = (torch.rand((wc, K), generator=gen, device=dev, dtype=torch.float64) * (spare[:, None] + 1).to(torch.float64)).to(torch.int64)
        cuts = torch.minimum(cuts, spare[:, None]).sort(dim=1).values
        edges = torch.cat([torch.zeros((wc, 1), dtype=torch.int64, device=dev), cuts, spare[:, None]], dim=1)
        bb = edges[:, 1:] - edges[:, :-1]                       # (wc, K + 1), sums to spare
        nl = node_len[w0:w0 + wc]
        nl[:, 0] = bb[:, 0].to(torch.int32)
        nl[:, ref_col] = ref_len.to(torch.int32)
        nl[:, ref_col + 1] = alt_len.to(torch.int32)
        nl[:, ref_col + 2] = bb[:, 1:].to(torch.int32)
        k_anc = torch.multinomial(sfs, wc * K, replacement=True, generator=gen).reshape(wc, K) + 1
        p_anc = (k_anc.to(torch.float64) / n).clamp(1e-6, 1 - 1e-6)
        ga = torch._standard_gamma((p_anc[:, None, :] * a_bn).expand(wc, npop, K).contiguous(), generator=gen)
        gb = torch._standard_gamma(((1.0 - p_anc[:, None, :]) * a_bn).expand(wc, npop, K).contiguous(), generator=gen)
        p_pop = (ga / (ga + gb).clamp_min(1e-300)).to(torch.float32)          # Beta(a, b) = Ga / (Ga + Gb)
        u = torch.rand((wc, n, K), generator=gen, device=dev)
        alt = u < p_pop[:, pops_t, :]
        dense = torch.zeros((wc, n, m_pad), dtype=torch.uint8, device=dev)
        dense[:, :, 0] = 1
        dense[:, :, ref_col] = (~alt).to(torch.uint8)
        dense[:, :, ref_col + 1] = alt.to(torch.uint8)
        dense[:, :, ref_col + 2] = 1
        x_bits[w0:w0 + wc] = ctx.pack_bits(dense.view(wc * n, m_pad), pitch).view(wc, n, pitch)
        del dense, u, alt
    ctx.check()
    return x_bits, node_len, pops, m, m_pad


class HostGenerator:
    """Stand-in for a device Context so make_windows_device can run on the host (CPU torch + numpy
    packbits): used where no GPU work is wanted, e.g. the CPU reference arm of bench.py."""

    def __init__(self):
        import torch
        self.torch_device = torch.device("cpu")

    def pack_bits(self, dense, pitch):
        import torch
        d = dense.numpy()
        packed = np.packbits(d, axis=1, bitorder="little")
        return torch.from_numpy(packed.view("<u4").view(np.int32).reshape(d.shape[0], pitch))

    def check(self):
        pass
