"""View-sharded data parallelism: Gaussians and EMD state are replicated, every
rank renders its own (timestep, cameras) views, and the parameter gradients are summed across ranks.
The reference itself is single-GPU (``assert batch_size == 1``, OmniRe/models/trainers/base.py:411; no
``init_process_group`` on the path), so this is new plumbing around the unchanged per-rank step.

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on B200; gloo in the CPU tests).
The only exchange of the path is the gradient all-reduce, so there is no custom collective kernel:
the compute step is not followed by a collective it could be fused with tile by tile (raster backward
produces per-(tile,Gaussian) partials that are reduced on-device first; what crosses NVLink is the
already-reduced dense gradient).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(num_timesteps: int, rank: int, world: int) -> List[int]:
    """Timesteps rendered by ``rank``: round-robin, so consecutive frames land on different GPUs."""
    return list(range(rank, num_timesteps, world))


def allreduce_grads(params: Iterable[Tensor], average: bool = False, bucket_bytes: int = 64 << 20) -> int:
    """Sum (or average) ``.grad`` of every parameter across ranks.

    Small tensors (track heads, pose rows, temporal tables) are packed into flat buckets so the launch
    count stays low; large per-Gaussian tensors go out as they are.  All operations are issued
    asynchronously and waited for once.  Returns the number of bytes reduced (per rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    big, small = [], []
    for p in params:
        if p.grad is None:
            continue
        (big if p.grad.numel() * p.grad.element_size() >= (1 << 20) else small).append(p.grad)
    works, total = [], 0
    for g in big:
        works.append(dist.all_reduce(g, async_op=True))
        total += g.numel() * g.element_size()
    flat_sets: List[Tuple[Tensor, List[Tensor]]] = []
    cur, cur_bytes = [], 0
    for g in small:
        nbytes = g.numel() * g.element_size()
        if cur and cur_bytes + nbytes > bucket_bytes:
            flat_sets.append((torch.cat([x.reshape(-1) for x in cur]), cur))
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    if cur:
        flat_sets.append((torch.cat([x.reshape(-1) for x in cur]), cur))
    for flat, _ in flat_sets:
        works.append(dist.all_reduce(flat, async_op=True))
        total += flat.numel() * flat.element_size()
    for w in works:
        w.wait()
    for flat, gs in flat_sets:
        off = 0
        for g in gs:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    if average:
        for p in params:
            if p.grad is not None:
                p.grad.div_(world)
    return total


def allreduce_densify_stats(grad_norm_sum: Tensor, vis_count: Tensor, max_radii: Optional[Tensor] = None) -> None:
    """Densification statistics are not gradients but must agree across replicas or they diverge at the next
    refinement (vanilla.py:171-191): sums for the accumulated |grad| and visibility counts, max for radii."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dist.all_reduce(grad_norm_sum)
    dist.all_reduce(vis_count)
    if max_radii is not None:
        dist.all_reduce(max_radii, op=dist.ReduceOp.MAX)
