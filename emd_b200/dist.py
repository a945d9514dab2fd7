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


def _flag_device(params) -> torch.device:
    for p in params:
        return p.device
    return torch.device("cpu")


def allreduce_grads(params: Iterable[Tensor], average: bool = False, bucket_bytes: int = 64 << 20) -> int:
    """Sum (or average) ``.grad`` of every parameter across ranks.

    Small tensors (track heads, pose rows, temporal tables) are packed into flat buckets so the launch
    count stays low; large per-Gaussian tensors go out as they are.  All operations are issued
    asynchronously and waited for once.  Returns the number of bytes reduced (per rank)."""
    params = list(params)   # may be a generator (``model.parameters()``): it is walked more than once
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    # every rank must issue the same sequence of collectives: a parameter without a gradient on THIS rank (a node class
    # none of this rank's views saw) takes part with zeros -- unless no rank has one, in which case it stays None
    used = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=_flag_device(params))
    if used.numel():
        dist.all_reduce(used, op=dist.ReduceOp.MAX)
    used = used.tolist()
    big, small = [], []
    for p, u in zip(params, used):
        if not u:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
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


def _allreduce_group_async(tensors: List[Tensor], group=None):
    """One asynchronous all-reduce over several tensors: a single grouped NCCL launch (``ncclGroupStart/End``
    around one ``ncclAllReduce`` per tensor -- no packing copies), gloo's coalesced all-reduce in the CPU tests."""
    if len(tensors) == 1:
        return dist.all_reduce(tensors[0], async_op=True, group=group)
    from torch.distributed.distributed_c10d import _coalescing_manager
    with _coalescing_manager(group=group, async_ops=True) as cm:
        for t in tensors:
            dist.all_reduce(t, group=group)
    return cm


class GradReducer:
    """Gradient all-reduce that starts while the backward pass is still running.

    ``early`` is a list of GROUPS of parameters that take part in EVERY step on EVERY rank (background SH
    coefficients first of all: 180 of the 252 B/Gaussian the exchange moves).  Every member gets a post-accumulate
    hook; the moment autograd has written the ``.grad`` of a group's last member, ONE grouped asynchronous
    all-reduce of the group is issued on the collective's own stream, and the rest of the backward pass
    (projection, EMD deformation) runs under it.
    ``finish()`` -- called once after ``backward()`` -- reduces everything else in one more grouped launch
    (a parameter that got no gradient on this rank, e.g. SMPL nodes of a frame without visible pedestrians,
    contributes zeros so every rank issues the same sequence; one that got none on ANY rank gets its ``.grad`` reset
    to None afterwards, so the optimizer skips it as the reference's does) and waits for all of it.
    Contract: exactly ONE ``backward()`` per ``finish()`` (the hooks reduce a gradient the moment it is first
    written; a second accumulation into it would race the all-reduce in flight) -- asserted.

    The early sequence is the order in which autograd completes the groups, which is the same on every rank as
    long as the early parameters are used by every rank's step; that is the contract of ``early``."""

    def __init__(self, params: Sequence[Tensor], early: Optional[Sequence[Sequence[Tensor]]] = None,
                 average: bool = False, pack_below: int = 4 << 20, drop_unused: bool = False, defer_early: bool = False,
                 tail_group=None):
        """``drop_unused``: exchange one flag per parameter so that a parameter no rank used keeps ``grad None`` (costs a
        host read of the flags per step; off, such a parameter gets a zero gradient).
        ``defer_early``: ``finish()`` returns without waiting for the EARLY groups' all-reduces; ``wait_deferred()``
        completes them.  The early groups are the SH coefficients, which the next step only reads after its projection
        and tile binning, so their exchange (and their optimizer update, which the caller runs after
        ``wait_deferred()``) hides behind the next step's front end.
        ``tail_group``: a second process group (``dist.new_group()``: its own communicator and stream) for what
        ``finish()`` reduces.  Collectives of one communicator run in issue order, so on the default group the tail would
        queue behind the early groups' all-reduce and waiting for the tail would wait for both -- with ``defer_early`` the
        tail needs its own communicator for the deferral to hide anything."""
        self.params = list(params)
        self.average = average
        self.pack_below = pack_below
        self.drop_unused = drop_unused
        self.defer_early = defer_early
        self.tail_group = tail_group
        self._deferred: List = []
        self._early_works: List = []
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.groups: List[List[Tensor]] = [list(g) for g in (early or []) if len(g)]
        self._group_of = {id(p): gi for gi, g in enumerate(self.groups) for p in g}
        self._pending = [len(g) for g in self.groups]
        self._seen = set()
        self._group_launched = [False] * len(self.groups)
        self._works: List = []
        self._handles = []
        self.bytes = 0
        self.early_bytes = 0
        if self.active:
            for g in self.groups:
                for p in g:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _launch(self, tensors: List[Tensor], early: bool = False) -> int:
        if early:
            self._early_works.append(_allreduce_group_async(tensors))
        else:
            self._works.append(_allreduce_group_async(tensors, group=self.tail_group))
        n = sum(t.numel() * t.element_size() for t in tensors)
        self.bytes += n
        return n

    def wait_deferred(self) -> None:
        """Complete the early groups' all-reduces left in flight by ``finish()`` (``defer_early``).  Call before anything
        reads or resets those gradients: the optimizer update of the early parameters, the next step's ``grad = None``."""
        if not self._deferred:
            return
        for w in self._deferred:
            w.wait()
        self._deferred = []
        if self.average:
            world = dist.get_world_size()
            for g in self.groups:
                for p in g:
                    if p.grad is not None:
                        p.grad.div_(world)

    def _hook(self, p: Tensor) -> None:
        if p.grad is None:
            return
        if self._deferred:
            raise RuntimeError("GradReducer: the previous step's deferred all-reduces are still in flight; call "
                               "wait_deferred() before the backward pass reaches the early parameters")
        if id(p) in self._seen:
            raise RuntimeError("GradReducer: a second backward() accumulated into an early-group gradient before finish(); "
                               "run one backward() per finish() (or construct the reducer without `early`)")
        self._seen.add(id(p))
        gi = self._group_of[id(p)]
        self._pending[gi] -= 1
        if self._pending[gi] == 0 and not self._group_launched[gi]:
            self._group_launched[gi] = True
            self.early_bytes += self._launch([q.grad for q in self.groups[gi]], early=True)

    def finish(self) -> int:
        """Reduce what the hooks did not, wait for everything; returns the bytes reduced this step (per rank)."""
        if not self.active:
            return 0
        world = dist.get_world_size()

        # which parameters got a gradient on ANY rank (one tiny MAX all-reduce, issued first on every rank)
        mine = [p for p in self.params if p.requires_grad] if self.drop_unused else []
        used_work = None
        if mine:
            used = torch.tensor([0.0 if p.grad is None else 1.0 for p in mine], device=_flag_device(mine))
            used_work = dist.all_reduce(used, op=dist.ReduceOp.MAX, async_op=True)

        def grad_of(p):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            return p.grad

        for gi, g in enumerate(self.groups):   # an early group whose hooks did not all fire (unused this step)
            if not self._group_launched[gi]:
                self._group_launched[gi] = True
                self._launch([grad_of(q) for q in g], early=True)
        # everything else: ONE grouped launch; tensors below `pack_below` bytes travel inside one flat buffer (a
        # collective per 100-byte tensor costs a full cross-GPU latency each: 31 tensors ~ 0.5 ms at 8 GPUs)
        rest = [grad_of(p) for p in self.params if p.requires_grad and id(p) not in self._group_of]
        big = [g for g in rest if g.numel() * g.element_size() >= self.pack_below]
        small = [g for g in rest if g.numel() * g.element_size() < self.pack_below]
        flat = None
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            big.append(flat)
        if big:
            self._launch(big)
        for w in self._works:
            w.wait()
        if self.defer_early:
            self._deferred, self._early_works = self._early_works, []
        else:
            for w in self._early_works:
                w.wait()
            self._early_works = []
        if flat is not None:
            torch._foreach_copy_(small, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in small]), small)])
        if used_work is not None:
            used_work.wait()
            for p, u in zip(mine, used.tolist()):
                if not u:
                    p.grad = None       # unused on every rank: stays without a gradient, like in the reference
        if self.average:
            for p in self.params:
                if p.grad is not None and not (self.defer_early and id(p) in self._group_of):
                    p.grad.div_(world)
        total = self.bytes
        self._works, self._seen, self.bytes, self.early_bytes = [], set(), 0, 0
        self._pending = [len(g) for g in self.groups]
        self._group_launched = [False] * len(self.groups)
        return total

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def allreduce_densify_stats(grad_norm_sum: Tensor, vis_count: Tensor, max_radii: Optional[Tensor] = None) -> None:
    """Densification statistics are not gradients but must agree across replicas or they diverge at the next
    refinement (vanilla.py:171-191): sums for the accumulated |grad| and visibility counts, max for radii."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dist.all_reduce(grad_norm_sum)
    dist.all_reduce(vis_count)
    if max_radii is not None:
        dist.all_reduce(max_radii, op=dist.ReduceOp.MAX)
