"""Parity of the BENCHMARKED plans against the oracle.

The production planner takes decisions at batch 32 that the small whole-net tests never trigger: the CTA-pair
kernel auto-selected for res4, MT=2 persistent tiles, the wave-aware N-tile split (res5: Cout 512 as 3 x 176
columns with a ragged last tile), 3-D depth 16 -> 8 -> 4.  Videos are independent in TEST phase (proven bitwise
by test_eco_lite_n16_properties), so single videos of the batch are checked against the oracle run on that clip
alone (B=1, about a second each)."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import (TOL_LOGITS, TOL_NET, TOL_NET_FULL, check_bf16_blob, describe_mismatch, load_params, rel_l2,
                         rel_max, teacher_blobs, teacher_raw_blobs)
from test_gpu_ops import two_conv_net

pytestmark = pytest.mark.gpu


def production_net(txt, keep_all=False, graph=True, **extra):
    """exactly what bench.py builds: default planner options, nothing forced"""
    import caffe
    opts = {"keep_all_blobs": 1 if keep_all else 0, "use_graph": 1 if graph else 0}
    opts.update(extra)
    return caffe.Net.from_string(txt, caffe.TEST, **opts)


def oracle_for(make, segments, classes):
    ref = refnet.RefNet(make(segments=segments, classes=classes, batch=1)).init_params(4321)
    ref.calibrate_bn(refnet.eco_input(1, segments))
    return ref


def clip(x, v, segments):
    return x[v * segments:(v + 1) * segments]


@pytest.mark.parametrize("model", ["lite", "full"])
def test_benchmarked_plan_logits_vs_oracle(gpu, model):
    # BASELINE config #2 (ECO-Lite N=16, B=32, graph on, fast plan) and config #3's per-GPU geometry (ECO-Full N=16)
    segments, batch = 16, 32
    make = gen.eco_lite_deploy if model == "lite" else gen.eco_full_deploy
    classes = 101 if model == "lite" else 400
    ref = oracle_for(make, segments, classes)
    net = production_net(make(segments=segments, classes=classes, batch=batch))
    load_params(net, ref.params_dict())
    x = refnet.eco_input(batch, segments)
    for _ in range(2):  # second pass = CUDA-graph replay
        net.blobs["data"].data[...] = x
        got = net.forward()["fc8"].copy()
    assert got.shape == (batch, classes)
    tol = TOL_LOGITS if model == "lite" else 1.5 * TOL_LOGITS
    worst = 0.0
    for v in (0, 15, 31):
        want = ref.forward(clip(x, v, segments), bf16=True)["fc8"]
        e = rel_max(got[v:v + 1], want)
        worst = max(worst, e)
        assert e <= tol, describe_mismatch(got[v:v + 1], want, "fc8[video %d]" % v)
    print("benchmarked plan %s: worst logits rel_max %.3e" % (model, worst))


def test_benchmarked_plan_blobs_vs_oracle(gpu):
    # the same batch-32 planner decisions with every blob materialised: landmarks through trunk and head
    segments, batch = 16, 32
    ref = oracle_for(gen.eco_lite_deploy, segments, 101)
    net = production_net(gen.eco_lite_deploy(segments=segments, classes=101, batch=batch), keep_all=True, graph=False)
    load_params(net, ref.params_dict())
    x = refnet.eco_input(batch, segments)
    net.blobs["data"].data[...] = x
    net.forward()
    names = ["pool1_3x3_s2", "inception_3c_double_3x3_1_bn", "res3b_bn", "res4b_bn", "res5b_bn", "global_pool", "fc8"]
    dev = {n: net.blobs[n].data for n in names}
    for v in (0, 31):
        want = ref.forward(clip(x, v, segments), bf16=True)
        for n in names:
            w = want[n]
            per_video = w.shape[0]  # 16 frames in the trunk, 1 video in the head
            g = dev[n][v * per_video:(v + 1) * per_video]
            assert g.shape == w.shape, (n, g.shape, w.shape)
            e = min(rel_l2(g, w), rel_l2(g, refnet.round_bf16(w)))
            assert e <= TOL_NET, describe_mismatch(g, w, "%s[video %d]" % (n, v))


def test_online_plan_n8_b1_vs_oracle(gpu):
    # BASELINE config #5: sliding window N=8, one clip
    segments = 8
    ref = oracle_for(gen.eco_lite_deploy, segments, 101)
    net = production_net(gen.eco_lite_deploy(segments=segments, classes=101, batch=1))
    load_params(net, ref.params_dict())
    x = refnet.eco_input(1, segments)
    net.blobs["data"].data[...] = x
    got = net.forward()["fc8"].copy()
    want = ref.forward(x, bf16=True)["fc8"]
    assert rel_max(got, want) <= TOL_LOGITS, describe_mismatch(got, want, "fc8")


# ---- the exact production tiles, one layer at a time, auto planner (nothing forced), teacher-forced -> 1 bf16 ulp ----
PRODUCTION_TILES = [
    # res5 at batch 32: M = 32*4*7*7 = 6272, Cout 512 -> 3 N tiles of 176 columns (ragged last tile), grid 147
    ("res5_3x176", (32, 8, 4, 7, 7), 256, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1], dict(kernel="persistent", block_n=176, tiles=147)),
    # res4 at batch 32: M = 50176, 256-wide tiles, >= 2 tiles per SM pair -> conv_umma_pair_kernel by the cost model
    ("res4_auto_pair", (32, 8, 8, 14, 14), 256, 256, [3, 3, 3], [1, 1, 1], [1, 1, 1], dict(kernel="pair", block_n=256)),
    # res3 at batch 8: M = 100352, Cout 128 -> MT=2 persistent tiles (>= 2 tiles of 256 rows per SM)
    ("res3_mt2", (8, 8, 16, 28, 28), 128, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1], dict(kernel="persistent", block_n=128, mt=2)),
    # res3a_2n geometry: Cin 96 (64 + 32-channel tail block), MT=2
    ("res3a_cin96_mt2", (8, 8, 16, 28, 28), 96, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1], dict(kernel="persistent", mt=2, Cin=96)),
    # inception 3x3 with N=64 at many tiles per SM (MT=2, short N)
    ("inc3x3_n64", (256, 8, 28, 28), 64, 64, [3, 3], [1, 1], [1, 1], dict(kernel="persistent", block_n=64, mt=2)),
]


@pytest.mark.parametrize("case", PRODUCTION_TILES, ids=[c[0] for c in PRODUCTION_TILES])
def test_production_tile_shapes(gpu, case):
    _, shape, cmid, cout, k, s, p, expect = case
    txt = two_conv_net(shape, cmid, cout, k, s, p)
    ref = refnet.RefNet(txt).init_params(11)
    rng = np.random.default_rng(5)
    x = rng.normal(size=shape).astype(np.float32)
    net = production_net(txt, keep_all=True, graph=False)
    load_params(net, ref.params_dict())
    plan = {op["name"]: op for op in net.describe_plan()}
    for key, val in expect.items():
        assert plan["c"][key] == val, ("planner chose something else than this test is about", plan["c"])
    net.blobs["data"].data[...] = x
    net.forward()
    dev = {n: net.blobs[n].data.copy() for n in ("a_bn", "c", "c_bn")}
    forced = ref.forward(x, bf16=True, teacher=teacher_blobs(ref, dev), teacher_raw=teacher_raw_blobs(ref, dev))
    for n in ("c", "c_bn"):
        check_bf16_blob(dev[n], forced[n], n)
    # and the same layer in the fast plan (c fused away) gives the same c_bn
    net2 = production_net(txt, keep_all=False, graph=True)
    load_params(net2, ref.params_dict())
    net2.blobs["data"].data[...] = x
    net2.forward()
    assert {op["name"]: op for op in net2.describe_plan()}["c"]["kernel"] == expect["kernel"]
    dev2 = {"a_bn": net2.blobs["a_bn"].data.copy(), "c_bn": net2.blobs["c_bn"].data.copy()}
    forced2 = ref.forward(x, bf16=True, teacher=teacher_blobs(ref, dev2))
    check_bf16_blob(dev2["c_bn"], forced2["c_bn"], "c_bn (fast plan)")
