#!/usr/bin/env python
"""BASELINE config 5: online sliding-window recognition (scripts/online_recognition/online_recognition.py:64-97)
-- per-clip latency of `net.forward()` at batch 1 through the caffe-style Python surface, including the
host->device copy of the clip and the read-back of fc8.  Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import caffe
import gen_eco_prototxt as gen
from oracle import refnet
from eco_testlib import load_params

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, nargs="+", default=[8, 16])
ap.add_argument("--calls", type=int, default=200)
a = ap.parse_args()
caffe.set_device(0); caffe.set_mode_gpu()
ref = refnet.RefNet(gen.eco_lite_deploy(segments=4, classes=101, batch=1)).init_params(4321)
out = {"metric": "ECO-Lite online per-clip latency (B=1) through caffe.Net.forward, H2D + fc8 read-back included", "unit": "ms", "results": []}
for N in a.segments:
    net = caffe.Net.from_string(gen.eco_lite_deploy(segments=N, classes=101, batch=1), caffe.TEST, use_graph=1)
    load_params(net, ref.params_dict())
    x = refnet.eco_input(1, N)
    prev = np.zeros(101, np.float32)
    for _ in range(10):
        net.blobs["data"].data[...] = x
        net.forward()["fc8"]
    lat = []
    for _ in range(a.calls):
        t0 = time.perf_counter()
        net.blobs["data"].data[...] = x                 # online_recognition.py:92
        fc8 = net.forward()["fc8"][0]                   # :93
        prev = 0.5 * (prev + fc8)                       # running average with the previous prediction, :94-97
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat)
    out["results"].append({"segments": N, "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                           "mean_ms": float(lat.mean()), "clips_per_s": float(1e3 / lat.mean()),
                           "h2d_bytes": int(x.nbytes), "launches": net.last_launch_count()})
print(json.dumps(out))
