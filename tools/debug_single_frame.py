"""debug: per-blob comparison of a 1-frame net against frame 0 of an N-frame net (both keep_all / fast plans)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import caffe, gen_eco_prototxt as gen, harness

for keep in (1, 0):
    txtN = gen.eco_lite_deploy(segments=8, classes=101, batch=1)
    big = caffe.Net.from_string(txtN, caffe.TEST, keep_all_blobs=keep)
    harness.init_params(big, 1)
    x = harness.synthetic_frames(1, 8)
    for k in (1, 2, 3):
        from caffe.online import _with_frames
        small = caffe.Net.from_string(_with_frames(txtN, k), caffe.TEST, until_blob="inception_3c_double_3x3_1_bn", keep_all_blobs=keep)
        small.share_with(big)
        big.blobs["data"].data[...] = x
        big.forward()
        small.blobs["data"].data[...] = x[:k]
        small.forward()
        bad = []
        for name in small._blob_names:
            if "_split_" in name:
                continue
            try:
                a = small.blobs[name].data
                b = big.blobs[name].data[:k]
            except RuntimeError:
                continue
            if a.shape == b.shape and not np.array_equal(a, b):
                bad.append((name, float(np.abs(a - b).max())))
        print("keep_all=%d k=%d first differing blobs: %s" % (keep, k, bad[:4]))
        if keep == 0:
            for op in small.describe_plan()[:3]:
                print("   ", {q: op[q] for q in ("name", "kernel", "tiles", "mt") if q in op})
