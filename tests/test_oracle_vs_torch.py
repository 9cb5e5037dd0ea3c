"""Second, independent pin for the oracle: every layer type on ECO's path and the whole
ECO-Lite graph against torch's CPU fp32 functional ops (oneDNN/MKL) -- a different
implementation by different authors.  SURVEY.md 8(c): whole-network outputs are not pinned
by any reference fixture, so two independent restatements must agree instead."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import refnet
from oracle.torch_ref import torch_eco_lite
import gen_eco_prototxt as gen

torch.set_num_threads(8)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("shape,cout,k,s,p", [
    ((4, 3, 32, 32), 16, [7, 7], [2, 2], [3, 3]),
    ((2, 64, 14, 14), 32, [1, 1], [1, 1], [0, 0]),
    ((2, 24, 14, 14), 40, [3, 3], [1, 1], [1, 1]),
    ((2, 24, 15, 15), 40, [3, 3], [2, 2], [1, 1]),
    ((1, 12, 4, 14, 14), 20, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ((2, 12, 8, 14, 14), 20, [3, 3, 3], [2, 2, 2], [1, 1, 1]),
])
def test_conv(shape, cout, k, s, p):
    rng = np.random.default_rng(0)
    x = rng.normal(size=shape).astype(np.float32)
    w = (rng.normal(size=[cout, shape[1]] + k) / np.sqrt(shape[1] * np.prod(k))).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    y = refnet.conv_forward(x, w, b, k, s, p)
    fn = F.conv2d if len(k) == 2 else F.conv3d
    yt = fn(t(x), t(w), t(b), stride=s, padding=p).numpy()
    assert y.shape == yt.shape
    assert np.abs(y - yt).max() < 1e-4


def test_maxpool_ceil_mode():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, 8, 112, 112)).astype(np.float32)
    y = refnet.pool_forward(x, [3, 3], [2, 2], [0, 0], "MAX")
    yt = F.max_pool2d(t(x), 3, 2, 0, ceil_mode=True).numpy()
    assert y.shape == yt.shape == (2, 8, 56, 56)
    assert np.array_equal(y, yt)


def test_avepool_pad_inclusive():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2, 8, 28, 28)).astype(np.float32)
    y = refnet.pool_forward(x, [3, 3], [1, 1], [1, 1], "AVE")
    yt = F.avg_pool2d(t(x), 3, 1, 1, count_include_pad=True).numpy()
    assert np.abs(y - yt).max() < 1e-6
    x5 = rng.normal(size=(2, 8, 4, 7, 7)).astype(np.float32)
    y5 = refnet.pool_forward(x5, [4, 7, 7], [1, 1, 1], [0, 0, 0], "AVE")
    assert np.abs(y5.ravel() - x5.mean((2, 3, 4)).ravel()).max() < 1e-6


@pytest.mark.parametrize("segments,batch", [(4, 2)])
def test_eco_lite_whole_net(segments, batch):
    net = refnet.RefNet(gen.eco_lite_deploy(segments=segments, classes=101, batch=batch))
    net.init_params(4321)
    x = refnet.eco_input(batch, segments)
    net.calibrate_bn(x)
    blobs = net.forward(x)
    with torch.no_grad():
        fc8, res5b = torch_eco_lite(net.params_dict(), t(x), segments)
    assert blobs["fc8"].shape == (batch, 101)
    rel = np.linalg.norm(blobs["res5b_bn"] - res5b) / np.linalg.norm(res5b)
    assert rel < 1e-4, rel
    assert np.abs(blobs["fc8"] - fc8).max() / np.abs(fc8).max() < 1e-4
    # activations stay O(1) with the calibrated harness weights
    for k in ("conv1_7x7_s2_bn", "inception_3b_output", "res3b_bn", "res5b_bn"):
        assert 0.05 < blobs[k].std() < 20, (k, blobs[k].std())
