"""Data-parallel gradient exchange for the training config (BASELINE config #4): one process per GPU, identical
replicas, disjoint video shards, ONE exchange per iteration -- the SUM all-reduce of the parameter gradients, then
1 / world (what caffe_3d does under MPI: net.cpp:670-702 per-blob host-staged MPI_Allreduce, solver.cpp:310-347 the scale).

Here the gradients of all layers live in one contiguous fp32 device arena (eco_net_grad_arena); it is cut into a few
buckets in layer order, and each bucket's ncclAllReduce (torch.distributed, NCCL over NVLink / NVSwitch) is launched on a
side stream the moment backward has enqueued the kernels that produce it (eco_net_set_grad_bucket_hook), so the exchange
of the last layers overlaps the backward pass of the first ones.  The 1 / world factor is folded into the solver's update
kernel.  BN statistics stay per replica (type "BN", not a synchronised BN), like the reference.

torch.distributed is plumbing only (process group, NCCL communicator, streams); nothing here computes."""
from __future__ import annotations

import ctypes as C

from . import _caffe
from ._caffe import check, lib


def bucket_ranges(slot_offsets, slot_counts, slot_layers, arena_count, nbuckets):
    """Pure-Python mirror of Net::grad_buckets (csrc/net_train.inc) for tests and for CPU (gloo) runs: equal shares of the
    arena cut at layer boundaries, bucket 0 = the last layers.  Returns [(offset, count)]."""
    n = len(slot_offsets)
    if n == 0 or nbuckets <= 0:
        return []
    target = (arena_count + nbuckets - 1) // nbuckets
    out = []
    end = arena_count
    s = n - 1
    while s >= 0:
        first = s
        while first > 0 and end - slot_offsets[first] < target:
            first -= 1
        while first > 0 and slot_layers[first - 1] == slot_layers[first]:
            first -= 1
        out.append((slot_offsets[first], end - slot_offsets[first]))
        end = slot_offsets[first]
        s = first - 1
    return out


def allreduce_buckets(flat, ranges, group=None, async_op=True):
    """SUM all-reduce of flat[offset:offset+count] per bucket with torch.distributed (any backend); returns the works."""
    import torch.distributed as dist
    works = []
    for off, cnt in ranges:
        works.append(dist.all_reduce(flat[off:off + cnt], op=dist.ReduceOp.SUM, group=group, async_op=async_op))
    return works


class _DevArena(object):
    """zero-copy torch view of a raw device allocation"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class GradExchange(object):
    """Overlapped bucketed all-reduce of a solver's gradient arena.

        ex = GradExchange(solver, nbuckets=3)          # torch.distributed (nccl) must be initialised
        solver.step(1)                                  # backward fires the bucket hooks, the update waits for them
    """

    def __init__(self, solver, nbuckets=3, group=None, overlap=True):
        """overlap=True (default) starts each bucket's all-reduce from inside backward (side stream, behind an event);
        overlap=False starts all buckets when backward has been enqueued completely.  Measured on 8 x B200
        (profiles/r02d_bench_train_n8_*.json): the whole 150 MB arena all-reduces in 0.43 ms over NVSwitch (NVLS), 1.3 % of
        a 33 ms step; overlapped 33.67 ms/step, exchanged after backward 33.95 ms/step, one GPU alone 33.08 ms/step."""
        self.overlap = bool(overlap)
        self.pending = []
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.solver, self.net, self.group = solver, solver.net, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the net must run on a stream torch knows about, so that events can order it against the NCCL stream
        self.compute = torch.cuda.Stream()
        self.net.set_stream(self.compute.cuda_stream)
        self.side = torch.cuda.Stream()
        self._bucket_cb = _caffe.GRAD_BUCKET_FN(self._on_bucket)
        check(lib().eco_net_set_grad_bucket_hook(self.net._h, int(nbuckets), self._bucket_cb, None))
        _, g, n = self.net.arenas()
        self.grad = torch.as_tensor(_DevArena(g, n), device=torch.device("cuda", torch.cuda.current_device()))
        self.works = []
        self.bytes_per_iter = 0
        solver.set_grad_sync(self._finish, self.world)

    def _on_bucket(self, user, bucket, offset, count):
        torch = self.torch
        if self.world <= 1:
            return
        if not self.overlap:
            self.pending.append((offset, count))
            return
        ev = torch.cuda.Event()
        ev.record(self.compute)              # everything that produces this bucket is enqueued before this point
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self.works.append(self.dist.all_reduce(self.grad[offset:offset + count], op=self.dist.ReduceOp.SUM,
                                                   group=self.group, async_op=True))
        self.bytes_per_iter += int(count) * 4

    def _finish(self):
        # called by the solver between backward and update: the update kernels must see the reduced gradients
        torch = self.torch
        if self.pending:
            ev = torch.cuda.Event()
            ev.record(self.compute)          # backward is enqueued completely
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                for offset, count in self.pending:
                    self.works.append(self.dist.all_reduce(self.grad[offset:offset + count], op=self.dist.ReduceOp.SUM,
                                                           group=self.group, async_op=True))
                    self.bytes_per_iter += int(count) * 4
            self.pending = []
        with torch.cuda.stream(self.side):
            for w in self.works:
                w.wait()
        self.works = []
        ev = torch.cuda.Event()
        ev.record(self.side)
        self.compute.wait_event(ev)
        self.last_bytes, self.bytes_per_iter = self.bytes_per_iter, 0

    def broadcast_params(self, src=0):
        """Solver::SyncData (solver.cpp:349-367): every replica starts from rank 0's weights"""
        if self.world <= 1:
            return
        p, _, n = self.net.arenas()
        t = self.torch.as_tensor(_DevArena(p, n), device=self.torch.device("cuda", self.torch.cuda.current_device()))
        self.torch.cuda.synchronize()
        self.dist.broadcast(t, src=src, group=self.group)
        self.torch.cuda.synchronize()
        self.net.params_updated_on_device()
