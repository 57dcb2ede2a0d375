"""Windows as column ranges of ONE chromosome-scale presence matrix.

The reference re-extracts a graph per BED row (`impg similarity -r REGION`, run_h-fst.sh:155-194): a sliding-window
scan re-reads the same alignments once per overlapping window.  Here the chromosome's haplotype x node matrix is
ingested once and stays in HBM; a window [start, end) is the run of nodes whose reference coordinate falls inside it,
described to libimpop_b200 as a column range -- no copy of the presence bits, whatever the overlap between windows:

    x_off       = first 128-node group of the range (16-byte aligned column start)
    pitch_words = pitch of the chromosome matrix
    m           = nodes from that group's start to the end of the range
    node_len    = the chromosome's node lengths over that span, ZERO for the nodes of the first group that precede
                  the window (a zero-length node adds nothing to I, A, U or S, so the leading bits are inert)

Only the per-window node-length vectors are materialised (4 bytes per node and window, built on the device with torch
index arithmetic: plumbing, no statistic is computed here).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import Context, WindowBatch

GROUP = 128          # nodes per 16-byte column group: the alignment libimpop_b200 needs for x_off


class ChromosomeMatrix:
    """Presence bits [n, pitch_words] (u32 words, row-major), node lengths [m] and the reference coordinate of
    every node (non-decreasing: nodes sorted along the reference; an off-reference node carries the coordinate of
    its bubble).  Arrays may be numpy (uploaded) or device tensors (used in place)."""

    def __init__(self, ctx: Context, x_bits, node_len, node_pos, names=None):
        self.ctx = ctx
        dev = ctx.torch_device
        xb = x_bits if isinstance(x_bits, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_bits, dtype=np.uint32).view(np.int32))
        self.n, self.pitch = int(xb.shape[0]), int(xb.shape[1])
        if self.pitch % 4:
            raise ValueError("pitch_words must be a multiple of 4 (16-byte rows)")
        self.x = xb.to(dev).contiguous()
        nl = node_len if isinstance(node_len, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(node_len, dtype=np.uint32).view(np.int32))
        self.node_len = nl.to(dev).contiguous()
        self.m = int(self.node_len.shape[0])
        if self.m > self.pitch * 32:
            raise ValueError("more nodes than columns")
        self.node_pos = np.ascontiguousarray(node_pos, dtype=np.int64)
        if self.node_pos.shape[0] != self.m or (np.diff(self.node_pos) < 0).any():
            raise ValueError("node_pos must give one non-decreasing coordinate per node")
        self.names = list(names) if names is not None else None

    def node_ranges(self, starts, ends):
        """[k0, k1) per window: nodes whose coordinate lies in [start, end)."""
        starts, ends = np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64)
        return np.searchsorted(self.node_pos, starts, "left"), np.searchsorted(self.node_pos, ends, "left")

    def windows(self, starts, ends, labels, stream=None) -> WindowBatch:
        """A WindowBatch over the BED rows (starts[i], ends[i]); labels: [n] uint8 shared by every window, or [W, n]."""
        k0, k1 = self.node_ranges(starts, ends)
        W = int(k0.shape[0])
        ka = (k0 // GROUP) * GROUP                         # aligned start of every window's column range
        m_w = np.maximum(k1 - ka, 0).astype(np.int64)
        m_w[k1 <= k0] = 0                                  # empty window: no nodes at all
        dev = self.ctx.torch_device
        len_off = np.zeros(W + 1, dtype=np.int64)
        np.cumsum(m_w, out=len_off[1:])
        total = int(len_off[-1])
        # per-window node lengths: gather node_len[ka + t], zero where ka + t < k0  (device-side index arithmetic)
        if total:
            w_of = torch.repeat_interleave(torch.arange(W, device=dev), torch.from_numpy(m_w).to(dev))
            t = torch.arange(total, device=dev) - torch.from_numpy(len_off[:-1]).to(dev)[w_of]
            src = torch.from_numpy(ka).to(dev)[w_of] + t
            lens = torch.where(src >= torch.from_numpy(k0).to(dev)[w_of], self.node_len[src], torch.zeros((), dtype=torch.int32, device=dev))
        else:
            lens = torch.zeros(1, dtype=torch.int32, device=dev)
        lab = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels, dtype=np.uint8))
        lab = lab.to(dev).contiguous()
        per_window = lab.dim() == 2
        L = np.asarray(ends, dtype=np.int64) - np.asarray(starts, dtype=np.int64)
        return WindowBatch(self.ctx, np.full(W, self.n), m_w, np.full(W, self.pitch), ka // 32, len_off[:-1],
                           np.arange(W, dtype=np.int64) * self.n if per_window else np.zeros(W, dtype=np.int64), L,
                           self.x, lens.contiguous(), lab, stream=stream)

    def sliding(self, length: int, step: int, labels, begin: int | None = None, end: int | None = None, stream=None):
        """(starts, ends, WindowBatch) for windows of `length` bp every `step` bp over [begin, end)."""
        lo = int(self.node_pos[0]) if begin is None and self.m else int(begin or 0)
        hi = int(self.node_pos[-1]) + 1 if end is None and self.m else int(end or 0)
        starts = np.arange(lo, max(hi - length, lo) + 1, step, dtype=np.int64)
        ends = starts + length
        return starts, ends, self.windows(starts, ends, labels, stream=stream)


def concatenate_windows(window_set):
    """Test / demo helper: the windows of a synth.WindowSet laid end to end as one chromosome
    -> (x_bits [n, pitch] uint32, node_len [m_total] uint32, node_pos [m_total] int64)."""
    W, n, m, L = window_set.windows, window_set.n, window_set.m, window_set.length
    dense = np.concatenate([window_set.dense(w)[:, :m] for w in range(W)], axis=1)
    node_len = np.concatenate([window_set.node_len[w, :m] for w in range(W)]).astype(np.uint32)
    m_total = W * m
    pitch = ((m_total + 127) // 128) * 4
    padded = np.zeros((n, pitch * 32), dtype=np.uint8)
    padded[:, :m_total] = dense
    x_bits = np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(n, pitch)
    # reference coordinate: window start + the node's rank within the window, clipped into the window
    pos = np.concatenate([w * L + np.minimum(np.arange(m, dtype=np.int64), L - 1) for w in range(W)])
    return x_bits, node_len, pos
