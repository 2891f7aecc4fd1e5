"""Multi-GPU plumbing of the path: one process per GPU, prompts sharded across ranks, no data-path collective.

The reference is pure data parallelism over prompts (Lightning DDP, per-rank seeds: launch.py:168,
custom/triplaneturbo/data/…multistep_v2.py:1072-1081).  The renderer needs no communication in forward; after the
backward the only exchange is the all-reduce of the trainable gradients that live on this path — the three decoder
MLPs (22 976 floats = 92 KB at C=32) and optionally the NeuS variance — issued as ONE flat buffer (latency-bound
over NVLink 5 / NVSwitch; NCCL picks NVLS when available).  ``dL/dspace_cache`` stays on the rank that owns the
prompt.  Works with any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests).
"""
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_prompts(n_prompts: int, rank: int, world_size: int) -> List[int]:
    """Prompt p is rendered by rank p mod world_size."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_prompts, world_size))


def local_batch(space_cache: torch.Tensor, rays: Sequence[torch.Tensor], views_per_prompt: int, rank: int,
                world_size: int) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Slice a global batch (space_cache [P,...], per-view tensors [P*V,...]) down to this rank's prompts."""
    mine = shard_prompts(space_cache.shape[0], rank, world_size)
    idx = torch.as_tensor(mine, device=space_cache.device, dtype=torch.long)
    vidx = (idx[:, None] * views_per_prompt + torch.arange(views_per_prompt, device=idx.device)[None]).reshape(-1)
    return space_cache.index_select(0, idx), [r.index_select(0, vidx) for r in rays]


def allreduce_gradients(grads: Iterable[torch.Tensor], average: bool = True, group=None) -> List[torch.Tensor]:
    """Sum (or average) the given gradient tensors over all ranks through one flat buffer; returns new tensors with
    the original shapes.  A no-op copy when torch.distributed is not initialised or the world has one rank."""
    grads = list(grads)
    if not grads:
        return []
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat = flat / dist.get_world_size(group)
    out, pos = [], 0
    for g in grads:
        out.append(flat[pos:pos + g.numel()].view_as(g))
        pos += g.numel()
    return out


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-side timing convention of bench.py: a step takes as long as its slowest rank."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
