#!/usr/bin/env python
"""Per-op time of one ECO-Lite training pass (forward ops, backward sub-steps) -> markdown table, sorted by time.
Usage: python tools/profile_train.py [--segments 32] [--batch 16] [--out profiles/r02_train_ops.md]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200")):
    sys.path.insert(0, p)
import numpy as np
import caffe, gen_eco_prototxt as gen, harness

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=32)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--out", default=None)
a = ap.parse_args()
net = caffe.Net.from_string(gen.eco_lite_train(segments=a.segments, classes=400, batch=a.batch), caffe.TRAIN)
harness.init_params(net, 4321)
net.blobs["data"].data[...] = harness.synthetic_frames(a.batch, a.segments).reshape(a.batch, 3 * a.segments, 224, 224)
net.blobs["label"].data[...] = np.random.default_rng(0).integers(0, 400, (a.batch, 1, 1, 1)).astype(np.float32)
for _ in range(2):
    net.forward(); net.clear_param_diffs(); net.backward()
rows = net.profile_train()
rows = net.profile_train()
tot = sum(r["ms"] for r in rows)
groups = {}
for r in rows:
    n = r["name"]
    key = "fwd conv" if n.startswith("fwd:") and any(n == "fwd:" + o["name"] for o in net.describe_plan() if o["type"] == "conv") else \
          "fwd other" if n.startswith("fwd:") else \
          "bwd wgrad" if n.endswith(":wgrad") else "bwd dgrad" if n.endswith(":dgrad") else "bwd bias" if n.endswith(":bias") else \
          "bwd residual" if n.endswith(":residual") else "bwd other"
    groups[key] = groups.get(key, 0.0) + r["ms"]
lines = ["# ECO-Lite N=%d B=%d training pass, per op (CUDA events, eager)" % (a.segments, a.batch), "",
         "total %.2f ms" % tot, "", "| group | ms | share |", "|---|---|---|"]
for k, v in sorted(groups.items(), key=lambda kv: -kv[1]):
    lines.append("| %s | %.3f | %.1f %% |" % (k, v, 100 * v / tot))
lines += ["", "| op | ms | share |", "|---|---|---|"]
for r in sorted(rows, key=lambda r: -r["ms"])[:60]:
    lines.append("| %s | %.3f | %.1f %% |" % (r["name"], r["ms"], 100 * r["ms"] / tot))
txt = "\n".join(lines)
print(txt)
if a.out:
    open(a.out, "w").write(txt + "\n")
