#!/usr/bin/env python
"""Development aid for GPU trips: run one tiny conv net in a given A mode and print where the
device result deviates from the oracle (per 8-row x 16-column block), so layout / descriptor bugs
are visible from a single log.  usage: debug_conv.py <a_mode> [case]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200"),
          os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

from oracle import refnet
from eco_testlib import load_params, make_net, rel_max
import test_gpu_ops as T

a_mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = sys.argv[2] if len(sys.argv) > 2 else "gemm"
if which == "gemm":
    shape, txt = (2, 16, 12, 12), T.two_conv_net((2, 16, 12, 12), 64, 64, [1, 1], [1, 1], [0, 0])
elif which == "c3x3":
    shape, txt = (2, 16, 12, 12), T.two_conv_net((2, 16, 12, 12), 64, 32, [3, 3], [1, 1], [1, 1])
elif which == "d3":
    shape, txt = (1, 8, 4, 10, 10), T.two_conv_net((1, 8, 4, 10, 10), 64, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1])
elif which == "stem":
    shape, txt = (2, 3, 32, 32), T.conv_net((2, 3, 32, 32), 64, [7, 7], [2, 2], [3, 3])
else:
    raise SystemExit("unknown case")
ref = refnet.RefNet(txt).init_params(1)
x = np.random.default_rng(0).normal(size=shape).astype(np.float32)
want = ref.forward(x, bf16=True)
net = make_net(txt, keep_all=True, a_mode=a_mode)
load_params(net, ref.params_dict())
net.blobs["data"].data[...] = x
net.forward()
net.sync()
for name in [n for n in ("a_bn", "c", "c_bn") if n in want]:
    g = net.blobs[name].data
    w = want[name]
    print("== %s a_mode=%d blob %s shape %s rel_max %.3e" % (which, a_mode, name, g.shape, rel_max(g, w)))
    if rel_max(g, w) > 2e-3:
        C = g.shape[1]
        gm = np.moveaxis(g, 1, -1).reshape(-1, C)   # rows = output positions (GEMM M), cols = channels (N)
        wm = np.moveaxis(w, 1, -1).reshape(-1, C)
        bad = np.abs(gm - wm) > 1e-2 * np.abs(wm).max()
        print("   bad fraction %.3f ; first rows:" % bad.mean())
        R = min(gm.shape[0], 144)
        for r0 in range(0, R, 8):
            print("   rows %3d-%3d " % (r0, r0 + 7) + " ".join("%2d" % bad[r0:r0 + 8, c0:c0 + 16].sum() for c0 in range(0, C, 16)))
        print("   got[0,:8]  ", np.round(gm[0, :8], 3))
        print("   want[0,:8] ", np.round(wm[0, :8], 3))
        print("   got[1,:8]  ", np.round(gm[1, :8], 3))
        print("   want[1,:8] ", np.round(wm[1, :8], 3))
        print("   zero frac got %.3f  nan %d" % ((gm == 0).mean(), np.isnan(gm).sum()))
print("debug_conv done")
