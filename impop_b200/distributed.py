"""Multi-GPU plumbing (SURVEY.md 8 e): one process per GPU, torch.distributed for the exchange.

Two partitionings, both without any data-path collective:
  * windows are independent -> contiguous blocks of windows per rank (`shard_bounds`), one
    all-gather of the per-window result rows at the end (`gather_rows`);
  * one large window (n ~ 10^4) -> its upper-triangular tile grid is dealt round-robin to the ranks
    (impop_window_sums rank/world), the [W, 4] partial sums are all-gathered and every rank adds
    them in rank order (impop_window_finalize), so the result is bit-identical on every rank and
    from run to run (`split_grid_stats`).
The collectives run on NCCL when the tensors are on the GPU and on gloo for the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, weights=None):
    """[world + 1] boundaries of contiguous shards.  Without weights: sizes differ by at most one.
    With per-unit weights (e.g. n^2 * m of each window): boundaries at equal cumulative weight."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if weights is None:
        base, extra = divmod(total, world)
        sizes = [base + (1 if r < extra else 0) for r in range(world)]
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    w = np.asarray(weights, dtype=np.float64)
    if w.shape[0] != total:
        raise ValueError("weights must have one entry per unit")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.clip(np.searchsorted(cum, targets, side="left"), 1, total)
    inner = np.where(targets - cum[inner - 1] < cum[inner] - targets, inner - 1, inner)   # nearest boundary
    return np.concatenate([[0], np.maximum.accumulate(inner), [total]]).astype(np.int64)


def shard_of(total: int, rank: int, world: int, weights=None):
    b = shard_bounds(total, world, weights)
    return int(b[rank]), int(b[rank + 1])


def gather_rows(local: torch.Tensor, bounds, group=None) -> torch.Tensor:
    """All-gather row blocks of unequal height into the full [total, cols] tensor (same on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    height = max(sizes) if sizes else 0
    cols = local.shape[1:]
    padded = torch.zeros((height, *cols), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    flat = torch.empty((world * height, *cols), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, padded, group=group)      # concatenated along dim 0 (gloo and nccl agree)
    out = flat.view(world, height, *cols)
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def gather_parts(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally shaped tensors into [world, ...] (rank order)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local.unsqueeze(0).contiguous()
    local = local.contiguous()
    flat = torch.empty((world * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, local, group=group)
    return flat.view(world, *local.shape)


def split_grid_stats(batch, algo: int = 0, group=None, stream=None):
    """Statistics of a replicated batch whose tile grid is split over the ranks of `group`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sums = batch.window_sums(rank, world, algo, stream=stream)
    parts = gather_parts(sums, group)
    return batch.finalize(parts, stream=stream)
