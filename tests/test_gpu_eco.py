"""Whole-network GPU parity: ECO-Lite / ECO-Full through the public surface against the oracle's
bf16-mirror forward on identical harness weights and synthetic frames (SURVEY.md 8(d) inputs),
plus size-independent properties at the full N=16 configuration."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import (TOL_LOGITS, TOL_NET, TOL_NET_FULL, check_bf16_blob, check_f32_blob, describe_mismatch, load_params,
                         make_net, rel_l2, rel_max, teacher_blobs, teacher_raw_blobs)

pytestmark = pytest.mark.gpu

# blobs whose conv output is consumed by a fused residual add: never stored
PLAIN_LITE = {"global_pool", "global_pool_reshape", "fc8", "global_pool2D", "reshape_fc_st2", "pool_fusion_st2",
              "pool_fusion_st2D", "global_pool_gn02_reshape"}
FUSED_AWAY = {"res3b_2", "res4a_down", "res4b_2", "res5a_down", "res5b_2"}


def oracle_lite(segments, batch, classes=101, full=False):
    txt = (gen.eco_full_deploy if full else gen.eco_lite_deploy)(segments=segments, classes=classes, batch=batch)
    ref = refnet.RefNet(txt).init_params(4321)
    x = refnet.eco_input(batch, segments)
    ref.calibrate_bn(x)
    return txt, ref, x


@pytest.mark.parametrize("a_mode", [0, 1], ids=["gather", "im2col"])
def test_eco_lite_n4_every_blob(gpu, a_mode):
    # BASELINE config 1 geometry (N=4), every blob of deploy.prototxt
    txt, ref, x = oracle_lite(4, 2)
    want = ref.forward(x, bf16=True)
    net = make_net(txt, keep_all=True, a_mode=a_mode)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    out = net.forward()
    assert list(out.keys()) == ["fc8"]
    blobs = net.blobs
    dev = {name: blobs[name].data.copy() for name in want if name not in FUSED_AWAY and name in blobs}
    # (1) teacher-forced: every blob of deploy.prototxt, each layer fed the device's own inputs
    forced = ref.forward(x, bf16=True, teacher=teacher_blobs(ref, dev), teacher_raw=teacher_raw_blobs(ref, dev))
    worst_flip = ("", 0.0)
    for name, g in dev.items():
        assert g.shape == want[name].shape, (name, g.shape, want[name].shape)
        if name in ("data",):
            continue
        if name in PLAIN_LITE:
            check_f32_blob(g, forced[name], name)
        else:
            f = check_bf16_blob(g, forced[name], name)
            if f > worst_flip[1]:
                worst_flip = (name, f)
    # (2) free-running: error growth through the whole net stays within the calibrated band
    worst = ("", 0.0)
    for name, g in dev.items():
        e = min(rel_l2(g, want[name]), rel_l2(g, refnet.round_bf16(want[name])))
        if e > worst[1]:
            worst = (name, e)
        assert e <= TOL_NET, describe_mismatch(g, want[name], name)
    print("worst flip fraction", worst_flip, "worst free-running rel_l2", worst)
    assert rel_max(out["fc8"], want["fc8"]) <= TOL_LOGITS, describe_mismatch(out["fc8"], want["fc8"], "fc8")


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cudagraph"])
def test_eco_lite_n4_fast_plan_matches(gpu, graph):
    # the production plan (nothing extra materialised, optionally CUDA-graph replay) gives the same logits
    txt, ref, x = oracle_lite(4, 2)
    want = ref.forward(x, bf16=True)["fc8"]
    net = make_net(txt, keep_all=False, graph=graph)
    net.set_option("persistent", 1)   # production default: per-layer choice between the two conv kernels
    load_params(net, ref.params_dict())
    for _ in range(3):
        net.blobs["data"].data[...] = x
        got = net.forward()["fc8"].copy()
        assert rel_max(got, want) <= TOL_LOGITS, describe_mismatch(got, want, "fc8")
    with pytest.raises(RuntimeError):
        net.blobs["conv1_7x7_s2"].data  # fused away in the fast plan: loud, not stale
    with pytest.raises(RuntimeError):
        net.blobs["conv1_7x7_s2_bn"].data  # the stem rows kernel pools on chip: only pool1 is stored
    full = ref.forward(x, bf16=True)
    check_bf16_blob(net.blobs["pool1_3x3_s2"].data.copy(), full["pool1_3x3_s2"], "pool1_3x3_s2")


def test_eco_full_n4(gpu):
    txt, ref, x = oracle_lite(4, 2, classes=400, full=True)
    want = ref.forward(x, bf16=True)
    net = make_net(txt, keep_all=True)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    out = net.forward()
    blobs = net.blobs
    dev = {name: blobs[name].data.copy() for name in want if name not in FUSED_AWAY and name in blobs and name != "data"}
    forced = ref.forward(x, bf16=True, teacher=teacher_blobs(ref, dev), teacher_raw=teacher_raw_blobs(ref, dev))
    for name, g in dev.items():
        if name in PLAIN_LITE:
            check_f32_blob(g, forced[name], name)
        else:
            check_bf16_blob(g, forced[name], name)
    for name in ("inception_3c_output", "inception_4e_output", "inception_5b_output", "global_pool2D",
                 "pool_fusion_st2D", "global_pool", "global_pool_gn02_reshape", "fc8"):
        g, w = dev[name], want[name]
        e = min(rel_l2(g, w), rel_l2(g, refnet.round_bf16(w)))
        assert e <= TOL_NET_FULL, describe_mismatch(g, w, name)
    assert out["fc8"].shape == (2, 400)
    assert rel_max(out["fc8"], want["fc8"]) <= 1.5 * TOL_LOGITS, describe_mismatch(out["fc8"], want["fc8"], "fc8")


def test_eco_lite_n16_properties(gpu):
    """Full-size configuration (N=16): properties that do not need the oracle at this size --
    determinism, per-video independence (TEST-phase BN couples nothing across videos), and
    batch reshaping at run time (Layer::Forward re-runs Reshape, layer.hpp:447)."""
    segments, batch = 16, 3
    txt = gen.eco_lite_deploy(segments=segments, classes=101, batch=batch)
    ref = refnet.RefNet(gen.eco_lite_deploy(segments=4, classes=101, batch=1)).init_params(4321)
    ref.calibrate_bn(refnet.eco_input(1, 4))
    net = make_net(txt, keep_all=False)
    load_params(net, ref.params_dict())  # same parameter shapes for any N
    x = refnet.eco_input(batch, segments)
    net.blobs["data"].data[...] = x
    a = net.forward()["fc8"].copy()
    net.blobs["data"].data[...] = x
    b = net.forward()["fc8"].copy()
    assert np.array_equal(a, b), "forward is not deterministic"
    perm = [2, 0, 1]
    xp = x.reshape(batch, segments, 3, 224, 224)[perm].reshape(x.shape)
    net.blobs["data"].data[...] = xp
    c = net.forward()["fc8"].copy()
    assert np.array_equal(c, a[perm]), "videos are not independent of their batch position"
    # shrink the batch at run time: blob.reshape + net.reshape, as test_net.cpp:2269-2325 does
    net.blobs["data"].reshape(segments, 3, 224, 224)
    net.reshape()
    net.blobs["data"].data[...] = x[:segments]
    d = net.forward()["fc8"].copy()
    assert d.shape == (1, 101)
    assert np.array_equal(d[0], a[0])
    assert np.isfinite(a).all() and a.std() > 1e-3


def test_pipelined_forward_matches_blocking_forward(gpu):
    """eco_net_forward_pipelined (copy of clip k+1 overlapping compute of clip k, two buffer pairs)
    returns exactly what the caffe-style blocking forward returns, for a stream of different clips."""
    import ctypes
    import caffe
    from caffe import _caffe
    segments, batch = 4, 2
    txt, ref, x = oracle_lite(segments, batch)
    for graph in (False, True):
        net = make_net(txt, keep_all=False, graph=graph)
        load_params(net, ref.params_dict())
        clips = [np.roll(x, k, axis=0) * (1.0 + 0.1 * k) for k in range(5)]
        want = []
        for c in clips:
            net.blobs["data"].data[...] = c
            want.append(net.forward()["fc8"].copy())
        count, ocount = x.size, batch * 101
        bufs = []
        for _ in range(2):
            pi, po = ctypes.c_void_p(), ctypes.c_void_p()
            _caffe.check(_caffe.lib().eco_host_alloc(ctypes.byref(pi), count * 4))
            _caffe.check(_caffe.lib().eco_host_alloc(ctypes.byref(po), ocount * 4))
            bufs.append((pi, po,
                         np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_float)), (count,)),
                         np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_float)), (ocount,))))
        tickets, got = [None, None], [None] * len(clips)
        owner = [None, None]
        for k, c in enumerate(clips):
            s_ = k & 1
            if tickets[s_] is not None:
                net.wait(tickets[s_])
                got[owner[s_]] = bufs[s_][3].copy().reshape(batch, 101)
            bufs[s_][2][:] = c.ravel()
            tickets[s_] = net.forward_pipelined(bufs[s_][0].value, count, bufs[s_][1].value, ocount)
            owner[s_] = k
        for s_ in (0, 1):
            if tickets[s_] is not None:
                net.wait(tickets[s_])
                got[owner[s_]] = bufs[s_][3].copy().reshape(batch, 101)
        for k in range(len(clips)):
            assert np.array_equal(got[k], want[k]), "clip %d (graph=%s)" % (k, graph)
        for pi, po, _, _ in bufs:
            _caffe.lib().eco_host_free(pi)
            _caffe.lib().eco_host_free(po)


def test_pipelined_uint8_frames_match_fp32_path(gpu):
    """Raw uint8 frames + on-device mean subtraction (eco_net_forward_pipelined_u8) give bit-identical logits to
    feeding `frame - mean` as fp32 through the caffe-style blob (what the reference's DataTransformer produces)."""
    import ctypes
    from caffe import _caffe
    segments, batch = 4, 2
    txt, ref, _ = oracle_lite(segments, batch)
    net = make_net(txt, keep_all=False, graph=True)
    load_params(net, ref.params_dict())
    rng = np.random.default_rng(7)
    mean = [104.0, 117.0, 123.0]
    count, ocount = batch * segments * 3 * 224 * 224, batch * 101
    pi, po = ctypes.c_void_p(), ctypes.c_void_p()
    _caffe.check(_caffe.lib().eco_host_alloc(ctypes.byref(pi), count))
    _caffe.check(_caffe.lib().eco_host_alloc(ctypes.byref(po), ocount * 4))
    hin = np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_ubyte)), (count,))
    hout = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_float)), (ocount,))
    for _ in range(3):
        u8 = rng.integers(0, 256, size=(batch * segments, 3, 224, 224), dtype=np.uint8)
        net.blobs["data"].data[...] = u8.astype(np.float32) - np.array(mean, np.float32).reshape(1, 3, 1, 1)
        want = net.forward()["fc8"].copy()
        hin[:] = u8.ravel()
        net.wait(net.forward_pipelined_u8(pi.value, count, mean, po.value, ocount))
        assert np.array_equal(hout.reshape(batch, 101), want)
    _caffe.lib().eco_host_free(pi)
    _caffe.lib().eco_host_free(po)


def test_chunked_blocking_forward_is_bit_identical(gpu):
    """One caffe-style forward() on a host-written input blob runs split into sub-batches (H2D of slice k+1 under the compute
    of slice k, option h2d_chunks); videos are independent and the kernels' K order does not depend on the batch size, so
    the logits are bit-identical to the unsplit forward."""
    segments, batch = 4, 8
    txt, ref, _ = oracle_lite(segments, 1)
    txt = gen.eco_lite_deploy(segments=segments, classes=101, batch=batch)
    x = refnet.eco_input(batch, segments)
    plain = make_net(txt, keep_all=False, graph=True, h2d_chunks=1)
    plain.set_option("persistent", 1)
    load_params(plain, ref.params_dict())
    plain.blobs["data"].data[...] = x
    want = plain.forward()["fc8"].copy()
    _ = plain.blobs["pool1_3x3_s2"].data          # unsplit: materialised blobs of the fast plan stay readable
    for chunks in (2, 4):
        net = make_net(txt, keep_all=False, graph=True, h2d_chunks=chunks)
        net.set_option("persistent", 1)
        load_params(net, ref.params_dict())
        for rep in range(3):                      # repeated calls, rewritten input, CUDA-graph replay in the sub-nets
            xi = np.roll(x, rep * segments, axis=0)
            net.blobs["data"].data[...] = xi
            got = net.forward()["fc8"].copy()
            assert np.array_equal(got, np.roll(want, rep, axis=0)), (chunks, rep)
        with pytest.raises(RuntimeError, match="sub-batches"):
            net.blobs["pool1_3x3_s2"].data
        # new weights reach the sub-nets
        net.params["fc8u"][1].data[...] += 1.0
        net.blobs["data"].data[...] = x
        assert np.allclose(net.forward()["fc8"], want + 1.0, atol=1e-5)
