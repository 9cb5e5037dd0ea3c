#!/usr/bin/env python
"""A/B comparison of planner/kernel options on the same GPU in the same process:
per-op CUDA-event times (Net.profile_forward) for ECO-Lite N=16 at one batch size.
usage: ab_bench.py [--batch 32] cfgA=key:val,key:val cfgB=..."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, caffe
import gen_eco_prototxt as gen
from oracle import refnet
from eco_testlib import load_params

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--model", default="lite")
ap.add_argument("cfgs", nargs="+")
a = ap.parse_args()
B, N = a.batch, 16
mk = gen.eco_full_deploy if a.model == "full" else gen.eco_lite_deploy
classes = 400 if a.model == "full" else 101
txt = mk(segments=N, classes=classes, batch=B)
ref = refnet.RefNet(mk(segments=4, classes=classes, batch=1)).init_params(4321)
frames = (torch.randint(0, 256, (B * N, 3, 224, 224), device="cuda", dtype=torch.uint8).float() - 110.0)
torch.cuda.synchronize()
results = {}
names = None
for cfg in a.cfgs:
    label, _, body = cfg.partition("=")
    opts = {"keep_all_blobs": 0, "use_graph": 0}
    for kv in filter(None, body.split(",")):
        k, v = kv.split(":")
        opts[k] = int(v)
    net = caffe.Net.from_string(txt, caffe.TEST, **opts)
    load_params(net, ref.params_dict())
    net.set_input_device("data", frames.data_ptr(), frames.numel())
    for _ in range(2):
        net._forward(0, len(net.layers) - 1)
    net.sync()
    acc = None
    for _ in range(a.iters):
        prof = net.profile_forward()
        ms = np.array([o["ms"] for o in prof])
        acc = ms if acc is None else np.minimum(acc, ms)
        names = [o["name"] for o in prof]
        kinds = [o["kind"] for o in prof]
        flops = [o["flops"] for o in prof]
    results[label] = {"ms": acc, "names": names, "kinds": kinds, "flops": flops}
    del net
labels = list(results)
order = []
for l in labels:  # union of op names, first appearance order (plans may fuse different ops)
    for n in results[l]["names"]:
        if n not in order:
            order.append(n)
print("%-34s" % "op" + "".join("%12s" % l for l in labels))
for n in order:
    row = ""
    for l in labels:
        r = results[l]
        row += "%12.1f" % (r["ms"][r["names"].index(n)] * 1e3) if n in r["names"] else "%12s" % "-"
    print("%-34s" % n[:34] + row)
print("%-34s" % "TOTAL us" + "".join("%12.1f" % (results[l]["ms"].sum() * 1e3) for l in labels))
def conv_tf(r):
    conv = [i for i, k in enumerate(r["kinds"]) if k == 0]
    return sum(r["flops"][i] for i in conv) / r["ms"][conv].sum() / 1e9
print("%-34s" % "conv TFLOP/s" + "".join("%12.1f" % conv_tf(results[l]) for l in labels))
print("%-34s" % "videos/s (sum of ops)" + "".join("%12.1f" % (B / results[l]["ms"].sum() * 1e3) for l in labels))
