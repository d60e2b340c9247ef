"""Multi-GPU sharding of the query path: rows (frames, window-middle frames, cached embeddings)
shard across ranks in contiguous blocks with no communication while embedding; only the per-shard top-k candidates
are exchanged -- ONE all-gather of a packed message per query batch -- and every rank runs the same deterministic
k-way merge.  The reference has no counterpart (single device).

Message of one rank (`msg_bytes(q, k)` bytes; the layout libb200clip.so's b200clip_*_nccl / b200clip_topk_merge_packed
use): int64 idx[q][k] (GLOBAL indices, -1 = empty) | float32 score[q][k] | zero padding to a multiple of 16 bytes.

Transports:
  * NCCL (GPUs): `B200CLIP.sim_topk_sharded` hands the process group's ncclComm_t to b200clip_sim_topk_nccl -- K4 writes
    into this rank's slot of the gather buffer, ncclAllGather runs in place on the compute stream, the merge kernel
    follows; no torch kernel, no host synchronisation.
  * anything else (gloo in the CPU tests of this host logic): `exchange_messages` = one
    torch.distributed.all_gather_into_tensor on a preallocated buffer."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
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


def msg_bytes(q: int, k: int) -> int:
    """== b200clip_topk_msg_bytes(q, k)."""
    return (q * k * 12 + 15) & ~15


def pack_message(scores: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scores fp32 [q,k], idx int64 [q,k] -> uint8 [msg_bytes(q,k)] (same device)."""
    q, k = scores.shape
    if out is None:
        out = torch.zeros(msg_bytes(q, k), dtype=torch.uint8, device=scores.device)
    out[: q * k * 8].view(torch.int64).copy_(idx.reshape(-1))
    out[q * k * 8: q * k * 12].view(torch.float32).copy_(scores.reshape(-1))
    return out


def unpack_messages(gathered: torch.Tensor, q: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """uint8 [g, msg_bytes] -> (scores fp32 [g,q,k], idx int64 [g,q,k]) (views of `gathered`)."""
    g = gathered.shape[0]
    idx = gathered[:, : q * k * 8].view(torch.int64).view(g, q, k)
    scores = gathered[:, q * k * 8: q * k * 12].view(torch.float32).view(g, q, k)
    return scores, idx


def exchange_messages(msg: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One collective: this rank's message uint8 [m] -> uint8 [world, m] in rank order (preallocated `out` is reused)."""
    rank, world = world_info(group)
    if out is None:
        out = torch.empty(world, msg.numel(), dtype=torch.uint8, device=msg.device)
    if world == 1:
        out[0].copy_(msg)
        return out
    dist.all_gather_into_tensor(out.view(-1), msg.contiguous(), group=group)
    return out


def allgather_candidates(scores: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores fp32 [Q,k], idx int64 [Q,k] (GLOBAL indices, -1 = empty) -> ([world,Q,k], [world,Q,k]) through one
    all-gather of the packed message."""
    rank, world = world_info(group)
    if world == 1:
        return scores.unsqueeze(0), idx.unsqueeze(0)
    q, k = scores.shape
    gathered = exchange_messages(pack_message(scores.float(), idx.to(torch.int64)), group)
    cs, ci = unpack_messages(gathered, q, k)
    return cs.contiguous(), ci.contiguous()


def nccl_comm_ptr(device: torch.device, group=None) -> int:
    """The ncclComm_t (as an int) PyTorch's NCCL process group uses for `device`, or 0 when the backend is not NCCL /
    the build does not expose it.  The communicator is created lazily by the first collective: a barrier forces it."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != "nccl":
        return 0
    try:
        pg = group if group is not None else dist.group.WORLD
        backend = pg._get_backend(torch.device(device))
        ptr = int(backend._comm_ptr())
        if ptr == 0:
            dist.barrier(group=group, device_ids=[torch.device(device).index])
            ptr = int(backend._comm_ptr())
        return ptr
    except (AttributeError, RuntimeError):
        return 0


def sharded_topk(local_topk: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]], n_total: int,
                 merge: Callable[[torch.Tensor, torch.Tensor], tuple], group=None):
    """Runs `local_topk(lo, hi)` on this rank's block (must return global indices), exchanges candidates and
    merges.  `merge` is the k-way merge (product: B200CLIP.topk_merge -> b200clip_topk_merge)."""
    rank, world = world_info(group)
    lo, hi = shard_range(n_total, rank, world)
    s, i = local_topk(lo, hi)
    cs, ci = allgather_candidates(s, i, group)
    return merge(cs, ci)


def merge_messages_reference(gathered: np.ndarray, q: int, k: int):
    """numpy statement of the merge every rank runs on the gathered messages (descending score, ties -> higher global
    index, -1 = empty): what b200clip_topk_merge_packed computes.  Used by the CPU tests of the message format."""
    g = gathered.shape[0]
    idx = gathered[:, : q * k * 8].copy().view(np.int64).reshape(g, q, k)
    sc = gathered[:, q * k * 8: q * k * 12].copy().view(np.float32).reshape(g, q, k)
    out_s = np.full((q, k), -np.inf, np.float32)
    out_i = np.full((q, k), -1, np.int64)
    for qq in range(q):
        cand = [(float(sc[r, qq, j]), int(idx[r, qq, j])) for r in range(g) for j in range(k) if idx[r, qq, j] >= 0]
        cand.sort(key=lambda t: (t[0], t[1]), reverse=True)
        for j, (s, i) in enumerate(cand[:k]):
            out_s[qq, j], out_i[qq, j] = s, i
    return out_s, out_i
