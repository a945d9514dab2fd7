#!/usr/bin/env python
"""NCCL all-reduce timing probe for the gradient sizes of the bench scene (run under torchrun)."""
import os, sys, json
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
res = {}
for name, nbytes in (("sh_rest_234MB", 1_300_000 * 45 * 4), ("all_352MB", 351_892_048), ("means_15.6MB", 1_300_000 * 12), ("1MB", 1 << 20), ("64KB", 1 << 16)):
    t = torch.randn(nbytes // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_reduce(t)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res[name] = {"ms": round(ms, 4), "algbw_GBs": round(nbytes / ms / 1e6, 1), "busbw_GBs": round(nbytes / ms / 1e6 * 2 * (world - 1) / world, 1)}
if rank == 0:
    print(json.dumps({"world": world, "nccl": torch.cuda.nccl.version(), "allreduce": res}))
dist.destroy_process_group()
