"""CPU checks of the parity machinery itself: teacher-forced oracle forward and the bf16-ulp criterion."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import bf16_ulp, check_bf16_blob, teacher_blobs


def test_teacher_forcing_is_identity_on_own_blobs():
    net = refnet.RefNet(gen.eco_lite_deploy(segments=4, classes=11, batch=1)).init_params(3)
    x = refnet.eco_input(1, 4)
    net.calibrate_bn(x)
    free = {k: v.copy() for k, v in net.forward(x, bf16=True).items()}
    forced = net.forward(x, bf16=True, teacher=teacher_blobs(net, free))
    for k in free:
        assert np.array_equal(forced[k], free[k]), k


def test_teacher_forcing_localises_an_error():
    net = refnet.RefNet(gen.eco_lite_deploy(segments=4, classes=11, batch=1)).init_params(3)
    x = refnet.eco_input(1, 4)
    net.calibrate_bn(x)
    free = {k: v.copy() for k, v in net.forward(x, bf16=True).items()}
    bad = dict(free)
    bad["inception_3a_output"] = free["inception_3a_output"] * 1.05   # a "device bug" in one blob
    forced = net.forward(x, bf16=True, teacher=teacher_blobs(net, bad))
    with pytest.raises(AssertionError):
        check_bf16_blob(bad["inception_3a_output"], forced["inception_3a_output"], "inception_3a_output")
    # downstream layers are judged on the (wrong) inputs the device saw, so they still pass individually
    bad_down = net.forward(x, bf16=True, teacher=None if False else teacher_blobs(net, bad))
    check_bf16_blob(refnet.round_bf16(bad_down["res3a_bn"]), bad_down["res3a_bn"], "res3a_bn")


def test_bf16_ulp_and_flip_criterion():
    v = np.array([1.0, 2.0, 3.0, 0.75, 100.0], np.float32)
    assert np.allclose(bf16_ulp(v), [2 ** -7, 2 ** -6, 2 ** -6, 2 ** -8, 0.5])
    rng = np.random.default_rng(0)
    w = refnet.round_bf16(rng.normal(size=10000).astype(np.float32))
    g = w.copy()
    idx = rng.choice(10000, 100, replace=False)      # 1% one-ulp flips: accepted
    g[idx] += bf16_ulp(w[idx]).astype(np.float32)
    check_bf16_blob(g, w, "flips")
    g2 = w.copy()
    g2[0] += 3 * bf16_ulp(w[:1]).astype(np.float32)[0] + 1e-3  # a 3-ulp error: rejected
    with pytest.raises(AssertionError):
        check_bf16_blob(g2, w, "bad")
    g3 = w + bf16_ulp(w).astype(np.float32)           # everything off by one ulp: rejected (not "rare flips")
    with pytest.raises(AssertionError):
        check_bf16_blob(g3, w, "all")
