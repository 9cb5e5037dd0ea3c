"""Tiny torch.distributed helpers shared by bench.py (NCCL, one rank per GPU) and the CPU tests (gloo).

ECO inference shards by video: ranks never exchange activations; the only
collectives are the barrier around the timed region and the MAX over ranks of the device time."""
from __future__ import annotations

import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard(total_videos, world, rank):
    """Contiguous shard [lo, hi) of `total_videos` for `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(int(total_videos), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Group(object):
    def __init__(self, backend, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = env_rank()
        self.device = device
        self.active = self.world > 1
        if self.active and not dist.is_initialized():
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend, **kw)

    def barrier(self):
        if self.device is not None and self.device.type == "cuda":
            self.torch.cuda.synchronize()
        if self.active:
            self.dist.barrier()
        if self.device is not None and self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        if not self.active:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value):
        if not self.active:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64,
                              device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.active and self.dist.is_initialized():
            self.dist.destroy_process_group()
