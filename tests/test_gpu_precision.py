"""The fp32-faithful execution mode (option precision=1): BASELINE config #1 (ECO-Lite N=4, one clip)
must match the oracle's PURE fp32 forward -- caffe_3d's arithmetic, no bf16 mirror -- within the north star's 1e-3
relative, on every blob of deploy.prototxt.  (The reference's own layer tests use 1e-4 for convolutions and 1e-3 for BN:
test_convolution_layer.cpp:300-345, test_bn_layer.cpp:44-87.)

How: every feature map is stored as three bf16 planes [hi | lo | hi] (hi = bf16(v), lo = bf16(v - hi)) and every weight
as [w_hi | w_hi | w_lo]; the tcgen05 GEMM over the tripled K axis accumulates hi*w_hi + lo*w_hi + hi*w_lo in fp32, i.e.
16 significant bits per operand; bias / BN / residual / ReLU stay fp32 in the epilogue, pooling runs in fp32."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import describe_mismatch, load_params, rel_l2, rel_max
from test_gpu_ops import two_conv_net

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-3   # the north star's bar: relative to the largest entry of the blob (rel_max) and in rel-L2


def precise_net(txt, keep_all=True, graph=False):
    import caffe
    return caffe.Net.from_string(txt, caffe.TEST, precision=1, keep_all_blobs=1 if keep_all else 0, use_graph=1 if graph else 0)


CASES = [
    ((2, 16, 12, 12), 64, 64, [3, 3], [1, 1], [1, 1]),
    ((2, 8, 15, 15), 96, 160, [3, 3], [2, 2], [1, 1]),      # 3 x 96 = 288 "channels": 4.5 K blocks
    ((1, 8, 9, 9), 64, 352, [1, 1], [1, 1], [0, 0]),
    ((1, 8, 4, 10, 10), 64, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ((2, 8, 8, 14, 14), 128, 256, [3, 3, 3], [2, 2, 2], [1, 1, 1]),
]


@pytest.mark.parametrize("case", CASES, ids=["p%d" % i for i in range(len(CASES))])
def test_conv_bn_relu_vs_fp32_oracle(gpu, case):
    shape, cmid, cout, k, s, p = case
    txt = two_conv_net(shape, cmid, cout, k, s, p)
    ref = refnet.RefNet(txt).init_params(1)
    x = np.random.default_rng(0).normal(size=shape).astype(np.float32)   # NOT rounded to bf16
    want = ref.forward(x)                                                 # fp32 end to end
    net = precise_net(txt)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    net.forward()
    for name in ("a_bn", "c", "c_bn"):
        g = net.blobs[name].data
        assert rel_max(g, want[name]) <= 2e-4, describe_mismatch(g, want[name], name)


def test_eco_lite_n4_every_blob_vs_fp32_oracle(gpu):
    segments, batch = 4, 1
    txt = gen.eco_lite_deploy(segments=segments, classes=101, batch=batch)
    ref = refnet.RefNet(txt).init_params(4321)
    x = refnet.eco_input(batch, segments)
    ref.calibrate_bn(x)
    want = ref.forward(x)   # pure fp32: the reference's arithmetic
    net = precise_net(txt)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    out = net.forward()
    worst = ("", 0.0, 0.0)
    checked = 0
    for name, w in want.items():
        if name == "data" or name not in net.blobs:
            continue
        try:
            g = net.blobs[name].data
        except RuntimeError:
            continue   # conv outputs folded into a residual add (DESIGN: never stored)
        em, el = rel_max(g, w), rel_l2(g, w)
        checked += 1
        if em > worst[1]:
            worst = (name, em, el)
        assert em <= TOL_FP32 and el <= TOL_FP32, describe_mismatch(g, w, name)
    print("precision=1, ECO-Lite N=4: %d blobs, worst rel_max %.3e (rel_l2 %.3e) at %s; fc8 rel_max %.3e" % (
        checked, worst[1], worst[2], worst[0], rel_max(out["fc8"], want["fc8"])))
    assert checked >= 70, checked   # every blob of deploy.prototxt except the five conv outputs folded into residual adds
    # the default bf16 plan on the same net, for the record (~1e-2)
    import caffe
    fast = caffe.Net.from_string(txt, caffe.TEST)
    load_params(fast, ref.params_dict())
    fast.blobs["data"].data[...] = x
    print("default bf16 plan vs fp32 oracle: fc8 rel_max %.3e" % rel_max(fast.forward()["fc8"], want["fc8"]))


def test_eco_lite_n16_precise_fast_plan_logits(gpu):
    # BASELINE config #2 geometry in the fp32-faithful mode, nothing extra materialised, CUDA-graph replay
    segments, batch = 16, 2
    txt = gen.eco_lite_deploy(segments=segments, classes=101, batch=batch)
    ref = refnet.RefNet(gen.eco_lite_deploy(segments=segments, classes=101, batch=1)).init_params(4321)
    x = refnet.eco_input(batch, segments)
    ref.calibrate_bn(x[:segments])
    net = precise_net(txt, keep_all=False, graph=True)
    load_params(net, ref.params_dict())
    for _ in range(2):
        net.blobs["data"].data[...] = x
        got = net.forward()["fc8"].copy()
    for v in range(batch):
        want = ref.forward(x[v * segments:(v + 1) * segments])["fc8"]
        assert rel_max(got[v:v + 1], want) <= TOL_FP32, describe_mismatch(got[v:v + 1], want, "fc8[%d]" % v)
