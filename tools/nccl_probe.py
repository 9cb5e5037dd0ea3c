#!/usr/bin/env python
"""torchrun probe: time torch.distributed.all_reduce (NCCL) of the gradient-arena payload on this box and print what NCCL
chose (run with NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH to see transports / NVLS)."""
import json, os, sys, time
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for mb in (1, 16, 50, 150):
    t = torch.ones(mb * 1024 * 1024 // 4, device="cuda")
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["%dMB" % mb] = {"ms": ms, "algbw_GBps": mb / 1024 / (ms / 1e3), "busbw_GBps": mb / 1024 / (ms / 1e3) * 2 * (world - 1) / world}
if rank == 0:
    print(json.dumps({"world": world, "allreduce": out}))
dist.destroy_process_group()
