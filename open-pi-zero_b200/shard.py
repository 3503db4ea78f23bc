"""Batch sharding of `infer_action` across the GPUs of one box.

Samples are independent and every GPU holds a full replica, so
the only multi-GPU logic is: split the observation batch into contiguous
per-rank slices, run the replica, gather the `[B_r, H, A]` action chunks.
No collective sits on the data path; `torch.distributed` is only used for the
optional gather (one all_gather of a few KB) and for benchmark barriers.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `batch` samples for `rank`
    (the first `batch % world_size` ranks get one extra sample)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(inputs: Dict[str, torch.Tensor], world_size: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slice every tensor argument of `infer_action` along the batch dimension."""
    batch = next(v for v in inputs.values() if torch.is_tensor(v)).shape[0]
    lo, hi = shard_bounds(batch, world_size, rank)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch else v)
            for k, v in inputs.items()}


def gather_actions(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather per-rank action chunks back into `[batch, H, A]` (rank order =
    sample order).  Ragged shards are padded to the largest shard."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)
