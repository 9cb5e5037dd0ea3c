#!/usr/bin/env python
"""Run under torchrun on >= 2 GPUs: the overlapped bucketed NCCL exchange of caffe/parallel.py gives every rank the SUM of
all ranks' local gradients (checked against a plain torch all-reduce of the local gradients), the solver then applies the
1/world average, and the replicas stay bit-identical in their weights after several steps.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_exchange_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import caffe
import gen_eco_prototxt as gen
import harness
from caffe.parallel import GradExchange, _DevArena

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
caffe.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
SOLVER = 'base_lr: 0.01 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 clip_gradients: 40 solver_type: NESTEROV'
net_txt = gen.eco_lite_train(segments=4, classes=20, batch=2).replace("dropout_ratio: 0.3", "dropout_ratio: 0")
solver = caffe.NesterovSolver(solver_text=SOLVER, net_text=net_txt)
net = solver.net
harness.init_params(net, 100 + rank)          # different initial weights per rank: broadcast_params must fix that
ex = GradExchange(solver, nbuckets=3)
ex.broadcast_params(0)
rng = np.random.default_rng(7 + rank)         # different data per rank
x = harness.synthetic_frames(2, 4, seed=50 + rank).reshape(2, 12, 224, 224)
lab = rng.integers(0, 20, size=(2, 1, 1, 1)).astype(np.float32)
net.blobs["data"].data[...] = x
net.blobs["label"].data[...] = lab
_, g, n = net.arenas()
grad = torch.as_tensor(_DevArena(g, n), device="cuda")
# (1) local gradients, then their sum by a plain all-reduce
net.forward()
net.clear_param_diffs()
saved_hook_world = ex.world
ex.world = 1                                   # hooks fire but do nothing
net.backward()
torch.cuda.synchronize()
local_grad = grad.clone()
want = local_grad.clone()
dist.all_reduce(want)
# (1b) reproducibility of the local pass, and a synchronous exchange (hooks replayed after a device sync)
net.forward()
net.clear_param_diffs()
net.backward()
torch.cuda.synchronize()
repro = float((grad - local_grad).abs().max() / local_grad.abs().max())
diag = {"local_repro_rel": repro, "buckets": []}
nb = C_int = None
import ctypes as C
from caffe import _caffe
nbk = C.c_int()
_caffe.check(_caffe.lib().eco_net_num_grad_buckets(net._h, C.byref(nbk)))
sync_grad = grad.clone()
for b in range(nbk.value):
    off, cnt = C.c_size_t(), C.c_size_t()
    _caffe.check(_caffe.lib().eco_net_grad_bucket(net._h, b, C.byref(off), C.byref(cnt)))
    dist.all_reduce(sync_grad[off.value:off.value + cnt.value])
    torch.cuda.synchronize()
    sl = slice(off.value, off.value + cnt.value)
    diag["buckets"].append({"bucket": b, "offset": off.value, "count": cnt.value,
                            "sync_vs_want": float((sync_grad[sl] - want[sl]).abs().max() / want.abs().max())})
# (2) the overlapped exchange on the same data
ex.world = saved_hook_world
net.forward()
net.clear_param_diffs()
net.backward()
ex._finish()
torch.cuda.synchronize()
err = float((grad - want).abs().max() / want.abs().max())
for d in diag["buckets"]:
    sl = slice(d["offset"], d["offset"] + d["count"])
    d["overlapped_vs_want"] = float((grad[sl] - want[sl]).abs().max() / want.abs().max())
    d["overlapped_vs_local"] = float((grad[sl] - local_grad[sl]).abs().max() / want.abs().max())
# (3) a few solver steps: replicas must stay identical
for it in range(4):
    net.blobs["data"].data[...] = x
    net.blobs["label"].data[...] = lab
    loss = solver.step(1)
p, _, _ = net.arenas()
w = torch.as_tensor(_DevArena(p, n), device="cuda").clone()
w0 = w.clone()
dist.broadcast(w0, src=0)
# BN running statistics (lr_mult 0) are per replica by design (type "BN", not a synchronised BN): compare learnable blobs
learn = torch.zeros(n, dtype=torch.bool, device="cuda")
for sl in net.param_slots():
    if sl["lr_mult"] != 0:
        learn[sl["offset"]:sl["offset"] + sl["count"]] = True
same = bool(torch.equal(w[learn], w0[learn]))
out = {"rank": rank, "world": world, "diag": diag, "exchange_rel_err": err, "replicas_identical": same, "loss": float(loss),
       "bytes_per_iter": ex.last_bytes}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "exchange_check_rank%d.json" % rank), "w") as f:
    json.dump(out, f)
print(json.dumps(out))
dist.destroy_process_group()
assert err < 1e-4 and same, out
