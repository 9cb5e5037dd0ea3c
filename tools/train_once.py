#!/usr/bin/env python
"""A few ECO-Lite training passes (forward + backward + Nesterov update) with nothing else around them: the command ncu wraps
for the training launch list / the --set full capture of wgrad_umma_kernel."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200")):
    sys.path.insert(0, p)
import numpy as np
import caffe, gen_eco_prototxt as gen, harness
ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=32)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--passes", type=int, default=2)
a = ap.parse_args()
s = caffe.NesterovSolver(solver_text='base_lr: 0.001 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 clip_gradients: 40 solver_type: NESTEROV',
                         net_text=gen.eco_lite_train(segments=a.segments, classes=400, batch=a.batch))
harness.init_params(s.net, 4321)
s.net.blobs["data"].data[...] = harness.synthetic_frames(a.batch, a.segments).reshape(a.batch, 3 * a.segments, 224, 224)
s.net.blobs["label"].data[...] = np.random.default_rng(0).integers(0, 400, (a.batch, 1, 1, 1)).astype(np.float32)
for _ in range(a.passes):
    print("loss", s.step(1))
