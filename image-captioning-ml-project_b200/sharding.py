"""Multi-GPU plumbing for the decode path: images are independent, so each rank decodes a contiguous shard
and the only exchange is one all-gather of the finished captions, lengths and scores.
One process per GPU, torch.distributed (NCCL on GPUs; gloo works for the CPU tests)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split; the first `num_images % world_size` ranks take one extra image."""
    base, rem = divmod(num_images, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_captions(local: Dict[str, torch.Tensor], num_images: int, group: Optional[dist.ProcessGroup] = None
                    ) -> Dict[str, torch.Tensor]:
    """all-gather {"tokens" int32 [b,T], "lengths" int32 [b], "scores" float [b]} from every rank into
    full-batch tensors in image order.  Shards may be uneven (padded to the largest shard for the collective)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(num_images, r, world) for r in range(world)]
    pad_to = max(e - s for s, e in sizes)
    out = {}
    for key, t in local.items():
        b = t.shape[0]
        if b < pad_to:
            t = torch.cat([t, t.new_zeros((pad_to - b,) + tuple(t.shape[1:]))], dim=0)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out[key] = torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)
    return out
