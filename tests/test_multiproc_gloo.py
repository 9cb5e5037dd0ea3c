"""The N>1 path on CPU: world_size-2 gloo runs of the helpers bench.py uses under torchrun, and the
reference arm's contract under torchrun (rank 0 prints one JSON line, the other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%r, "tools"))
import torch
from dist_util import Group, shard
g = Group("gloo")
lo, hi = shard(7, g.world, g.rank)
g.barrier()
mx = g.max_over_ranks(10.0 + g.rank)       # device time per rank -> the slowest rank defines the step
tot = g.sum_over_ranks(hi - lo)
# one file per rank: two ranks printing to the shared stdout can interleave inside a line
with open(os.path.join(os.environ["ECO_TEST_OUT"], "rank%%d.json" %% g.rank), "w") as f:
    json.dump({"rank": g.rank, "world": g.world, "shard": [lo, hi], "max": mx, "total": tot}, f)
g.close()
''' % ROOT


def _torchrun(args, timeout=240, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    env["OMP_NUM_THREADS"] = "2"
    import socket
    with socket.socket() as sk:   # a free port: parallel test runs must not collide on the rendezvous
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_gloo_world2_shard_barrier_max(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = _torchrun([str(script)], extra_env={"ECO_TEST_OUT": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [json.loads((tmp_path / ("rank%d.json" % k)).read_text()) for k in (0, 1)]
    assert sorted(x["rank"] for x in rows) == [0, 1]
    assert all(x["world"] == 2 and x["max"] == 11.0 and x["total"] == 7.0 for x in rows)
    assert sorted(tuple(x["shard"]) for x in rows) == [(0, 4), (4, 7)]


def test_reference_arm_under_torchrun_prints_once():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--segments", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
