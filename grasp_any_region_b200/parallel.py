"""Data-parallel plumbing for the 8xB200 box: one process per GPU, a full weight replica per rank,
independent (image, mask-set, prompt) requests sharded across ranks, and the path's single
collective -- an all-gather of the fixed-shape decoded-token buffer.  The reference
initialises NCCL and never uses it (demo/gar_with_mask.py:79-81); nothing else crosses GPUs."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_range(n_requests: int, rank: int, world: int):
    """Contiguous block partition; the first n % world ranks take one extra request."""
    base, extra = divmod(n_requests, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_tokens(token_lists: Sequence[Sequence[int]], rows: int, max_new_tokens: int, device) -> torch.Tensor:
    """(rows, max_new_tokens) int32, -1 padded; rows >= len(token_lists) so every rank sends the same shape."""
    out = torch.full((rows, max_new_tokens), -1, dtype=torch.int32, device=device)
    for i, t in enumerate(token_lists):
        if len(t):
            out[i, : len(t)] = torch.tensor(list(t), dtype=torch.int32, device=device)
    return out


def gather_tokens(local: torch.Tensor, n_requests: int) -> List[List[int]]:
    """All-gather the per-rank token buffers and return the captions' ids in global request order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [[t for t in row.tolist() if t >= 0] for row in local[:n_requests]]
    world = dist.get_world_size()
    buf = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, local.contiguous())
    out: List[List[int]] = []
    per = local.shape[0]
    for r in range(world):
        lo, hi = shard_range(n_requests, r, world)
        for i in range(hi - lo):
            out.append([t for t in buf[r * per + i].tolist() if t >= 0])
    return out
