"""Multi-GPU sharding of the query path: frames shard across ranks in contiguous blocks with no
communication while embedding; only the per-shard top-k candidates (k x (fp32 score, int64 global index) per
query) are exchanged with ONE all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests of this host logic), then
every rank runs the same deterministic k-way merge.  The reference has no counterpart (single device)."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns [r*ceil(n/world), min(n, (r+1)*ceil(n/world)))."""
    if world <= 0 or rank < 0 or rank >= world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allgather_candidates(scores: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores fp32 [Q,k], idx int64 [Q,k] (GLOBAL indices, -1 = empty) -> ([world,Q,k], [world,Q,k]).
    One collective: score bits and indices are packed into a single int64 [Q,k,2] message."""
    rank, world = world_info(group)
    if world == 1:
        return scores.unsqueeze(0), idx.unsqueeze(0)
    msg = torch.stack([scores.contiguous().view(torch.int32).to(torch.int64), idx.to(torch.int64)], dim=-1).contiguous()
    out = torch.empty((world,) + tuple(msg.shape), dtype=msg.dtype, device=msg.device)
    dist.all_gather(list(out.unbind(0)), msg, group=group)   # one collective; works on NCCL and gloo
    cs = out[..., 0].to(torch.int32).view(torch.float32)
    ci = out[..., 1]
    return cs.contiguous(), ci.contiguous()


def sharded_topk(local_topk: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]], n_total: int,
                 merge: Callable[[torch.Tensor, torch.Tensor], tuple], group=None):
    """Runs `local_topk(lo, hi)` on this rank's block (must return global indices), exchanges candidates and
    merges.  `merge` is the k-way merge (product: B200CLIP.topk_merge -> b200clip_topk_merge)."""
    rank, world = world_info(group)
    lo, hi = shard_range(n_total, rank, world)
    s, i = local_topk(lo, hi)
    cs, ci = allgather_candidates(s, i, group)
    return merge(cs, ci)
