"""The N>1 training path on CPU (world_size-2 gloo): the bucket cutting used for the gradient exchange and the
SUM-all-reduce + 1/world averaging of caffe/parallel.py, on host tensors (the NCCL / device-arena path is the same code
fed CUDA tensors; it runs under torchrun on the GPU box: tools/train_exchange_check.py)."""
import json
import os
import subprocess
import sys

from caffe.parallel import bucket_ranges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path[:0] = [os.path.join(%r, "tools"), os.path.join(%r, "eco-efficient-video-understanding_b200")]
import torch
import torch.distributed as dist
from dist_util import Group
from caffe.parallel import bucket_ranges, allreduce_buckets
g = Group("gloo")
offs, cnts, layers = [0, 64, 192, 256, 1280], [50, 128, 64, 1000, 10], [0, 0, 3, 5, 5]
total = 1344
ranges = bucket_ranges(offs, cnts, layers, total, 3)
flat = torch.arange(total, dtype=torch.float32) * (g.rank + 1)      # rank r holds (r + 1) * i
works = allreduce_buckets(flat, ranges, async_op=True)
for w in works:
    w.wait()
flat /= g.world                                                     # the 1 / world of solver.cpp:332-337
want = torch.arange(total, dtype=torch.float32) * (1 + 2) / 2
with open(os.path.join(os.environ["ECO_TEST_OUT"], "rank%%d.json" %% g.rank), "w") as f:
    json.dump({"rank": g.rank, "ranges": ranges, "ok": bool(torch.allclose(flat, want))}, f)
g.close()
''' % (ROOT, ROOT)


def test_bucket_ranges_cut_at_layer_boundaries():
    offs, cnts, layers = [0, 64, 192, 256, 1280], [50, 128, 64, 1000, 10], [0, 0, 3, 5, 5]
    r = bucket_ranges(offs, cnts, layers, 1344, 3)
    assert r[0] == (256, 1088)           # the last layer's two blobs stay together and already exceed a third
    assert sum(c for _, c in r) == 1344 and r[-1][0] == 0
    for (o0, c0), (o1, c1) in zip(r, r[1:]):
        assert o1 + c1 == o0
    assert bucket_ranges(offs, cnts, layers, 1344, 1) == [(0, 1344)]
    assert bucket_ranges([], [], [], 0, 3) == []


def test_gloo_world2_bucketed_allreduce(tmp_path):
    from test_multiproc_gloo import _torchrun
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = _torchrun([str(script)], extra_env={"ECO_TEST_OUT": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [json.loads((tmp_path / ("rank%d.json" % k)).read_text()) for k in (0, 1)]
    assert all(x["ok"] for x in rows)
    assert rows[0]["ranges"] == rows[1]["ranges"]
