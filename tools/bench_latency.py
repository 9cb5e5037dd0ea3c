#!/usr/bin/env python
"""BASELINE config 5: online sliding-window recognition (scripts/online_recognition/online_recognition.py:64-97) --
per-step latency at batch 1 through the caffe-style Python surface, host->device copy of the frames and read-back of fc8
included, p50 / p99 over --calls steps:
  uncached : what the reference does, net.forward() on all N frames of the window for every new frame
  cached   : caffe.online.SlidingWindowRecognizer, the 2-D trunk on the NEW frame only + the 3-D head
Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200")):
    sys.path.insert(0, p)
import numpy as np
import caffe
import gen_eco_prototxt as gen
import harness
from caffe.online import SlidingWindowRecognizer

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, nargs="+", default=[8, 16])
ap.add_argument("--calls", type=int, default=200)
a = ap.parse_args()
caffe.set_device(0); caffe.set_mode_gpu()
small = caffe.Net.from_string(gen.eco_lite_deploy(segments=4, classes=101, batch=1), caffe.TEST, keep_all_blobs=1)
harness.init_params(small, 4321)
harness.calibrate_bn_on_device(small, harness.synthetic_frames(1, 4))
out = {"metric": "ECO-Lite online per-step latency (B=1), H2D of the frames + fc8 read-back included", "unit": "ms", "results": []}


def stats(lat):
    lat = np.array(lat)
    return {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
            "steps_per_s": float(1e3 / lat.mean())}


for N in a.segments:
    txt = gen.eco_lite_deploy(segments=N, classes=101, batch=1)
    net = caffe.Net.from_string(txt, caffe.TEST, use_graph=1)
    harness.copy_params(net, small)
    video = harness.synthetic_frames(1, N + a.calls + 16, seed=5)
    for _ in range(10):
        net.blobs["data"].data[...] = video[:N]
        net.forward()["fc8"]
    lat, prev = [], np.zeros(101, np.float32)
    for i in range(a.calls):
        t0 = time.perf_counter()
        net.blobs["data"].data[...] = video[i:i + N]    # online_recognition.py:92
        fc8 = net.forward()["fc8"][0]                   # :93
        prev = 0.5 * (prev + fc8)                       # running average with the previous prediction, :94-97
        lat.append((time.perf_counter() - t0) * 1e3)
    r = {"segments": N, "uncached": dict(stats(lat), h2d_bytes=int(video[:N].nbytes), launches=net.last_launch_count())}
    rec = SlidingWindowRecognizer(txt, N, new_frames=1)
    rec.load_params_from(small)
    for i in range(N + 10):
        rec.push(video[i:i + 1])
    lat = []
    for i in range(N + 10, N + 10 + a.calls):
        t0 = time.perf_counter()
        fc8 = rec.push(video[i:i + 1])[0]
        prev = 0.5 * (prev + fc8)
        lat.append((time.perf_counter() - t0) * 1e3)
    r["cached"] = dict(stats(lat), h2d_bytes=int(video[:1].nbytes))
    # same window -> same logits (bitwise), checked on the last step
    net.blobs["data"].data[...] = video[N + 10 + a.calls - N:N + 10 + a.calls]
    r["cached_equals_uncached"] = bool(np.array_equal(net.forward()["fc8"][0], fc8))
    out["results"].append(r)
print(json.dumps(out))
