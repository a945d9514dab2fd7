"""CPU, world_size 2, gloo: the N > 1 host logic (view sharding, bucketed gradient all-reduce,
densification-statistics reduction) -- the data path itself has no collective besides this one."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emd_b200 import dist as D
        g = torch.Generator().manual_seed(100 + rank)
        shapes = [(300000, 3), (300000, 4), (7, 36), (3,), (150, 32), (1,), (12, 5, 4)]
        params = [torch.zeros(s, requires_grad=True) for s in shapes]
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        params.append(torch.zeros(5, requires_grad=True))  # no grad on this one
        local = [p.grad.clone() if p.grad is not None else None for p in params]
        nbytes = D.allreduce_grads(params)
        # reference: plain per-tensor all_reduce
        for p, l in zip(params, local):
            if l is None:
                assert p.grad is None
                continue
            ref = l.clone()
            dist.all_reduce(ref)
            assert torch.equal(p.grad, ref), "bucketed all-reduce differs from per-tensor all-reduce"
        assert nbytes == sum(l.numel() * 4 for l in local if l is not None)
        # average=True through a GENERATOR of parameters (walked twice inside), and a gradient missing on one rank only
        ps2 = [torch.zeros(4, requires_grad=True), torch.zeros(3, requires_grad=True), torch.zeros(2, requires_grad=True)]
        ps2[0].grad = torch.full((4,), float(rank + 1))
        if rank == 0:
            ps2[1].grad = torch.full((3,), 4.0)
        D.allreduce_grads((q for q in ps2), average=True)
        assert torch.equal(ps2[0].grad, torch.full((4,), 1.5)), "average=True was skipped for a generator input"
        assert torch.equal(ps2[1].grad, torch.full((3,), 2.0)), "a gradient present on one rank only must still be reduced"
        assert ps2[2].grad is None, "a parameter unused on every rank keeps grad None"
        a, b, c = torch.full((4,), float(rank + 1)), torch.full((4,), 2.0), torch.tensor([1.0 + rank, 5.0 - rank])
        D.allreduce_densify_stats(a, b, c)
        assert torch.equal(a, torch.full((4,), 3.0)) and torch.equal(b, torch.full((4,), 4.0))
        assert torch.equal(c, torch.tensor([2.0, 5.0]))
        # GradReducer: hook-driven early all-reduce during backward + the rest at finish(); a parameter that gets no
        # gradient on ONE rank only (rank 1 skips `skip`) must not desynchronise the sequence
        gr = torch.Generator().manual_seed(7)
        big = torch.randn(400000, 3, generator=gr).requires_grad_(True)      # early
        mid = torch.randn(300000, 1, generator=gr).requires_grad_(True)      # >= 1 MB, reduced at finish
        tiny = torch.randn(6, 6, generator=gr).requires_grad_(True)          # flat bucket
        skip = torch.randn(9, generator=gr).requires_grad_(True)             # no grad on rank 1
        never = torch.randn(4, generator=gr).requires_grad_(True)            # no grad on any rank
        ps = [big, mid, tiny, skip, never]
        red = D.GradReducer(ps, early=[[big], [mid, tiny]], drop_unused=True)
        for it in range(2):
            for q in ps:
                q.grad = None
            loss = (big * (rank + 1.0)).sum() + (mid * mid).sum() * (rank + 2.0) + (tiny.sum() * (it + 1.0))
            if rank == 0:
                loss = loss + (skip * 3.0).sum()
            loss.backward()
            if it == 0:
                assert red.early_bytes == (big.numel() + mid.numel() + tiny.numel()) * 4, "early groups were not reduced from their hooks"
            n = red.finish()
            assert n == sum(q.numel() * 4 for q in ps)
            assert never.grad is None, "a parameter no rank used must come out of finish() without a gradient"
            assert torch.equal(big.grad, torch.full_like(big, 3.0))
            assert torch.allclose(mid.grad, 2 * mid.detach() * 5.0)
            assert torch.equal(tiny.grad, torch.full_like(tiny, 2.0 * (it + 1)))
            assert torch.equal(skip.grad, torch.full_like(skip, 3.0))
        red.close()
        # deferred early groups: finish() leaves their all-reduce in flight, wait_deferred() completes it (and averages)
        red2 = D.GradReducer(ps, early=[[big]], average=True, defer_early=True, tail_group=dist.new_group())
        for it in range(2):
            red2.wait_deferred()           # start of a step: nothing may still be in flight when grads are reset
            for q in ps:
                q.grad = None
            ((big * (rank + 1.0)).sum() + (mid.sum() * 2.0) + tiny.sum() + skip.sum() + never.sum()).backward()
            red2.finish()
            assert torch.equal(mid.grad, torch.full_like(mid, 2.0)), "non-deferred gradients are complete after finish()"
            red2.wait_deferred()
            assert torch.equal(big.grad, torch.full_like(big, 1.5)), "deferred gradient after wait_deferred(): mean of 1 and 2"
        red2.close()
        views = D.shard_views(11, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, views)
        assert sorted(sum(gathered, [])) == list(range(11))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_single_process_is_noop():
    from emd_b200 import dist as D
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    assert D.allreduce_grads([p]) == 0 and torch.equal(p.grad, torch.ones(3))
