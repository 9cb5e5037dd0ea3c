"""Pins the oracle (oracle/ref_cpu.c) to the reference's own known-answer tests.

Golden vectors: tests/golden/reference_known_answers.json, extracted from
caffe_3d/src/caffe/test/test_pooling_layer.cpp by tests/golden/make_reference_known_answers.py.
Closed-form / property checks restate the reference tests cited on each function.
"""
import json
import os

import numpy as np
import pytest

from oracle import refnet

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_pooling_known_answers(case):
    # the reference fills num x channels copies of the same plane (num=2, channels=2)
    x1 = np.array(case["input"], np.float32).reshape([1, 1] + case["in_shape"])
    x = np.tile(x1, [2, 2] + [1] * len(case["in_shape"]))
    y = refnet.pool_forward(x, case["kernel"], case["stride"], case["pad"], case["method"])
    assert list(y.shape) == [2, 2] + case["out_shape"]
    exp = np.array(case["output"], np.float32).reshape(case["out_shape"])
    for n in range(2):
        for c in range(2):
            assert np.abs(y[n, c] - exp).max() <= case["tol"] + 1e-12


def test_pool_output_shapes_eco():
    # pooling_layer.cpp:131-147 ceil mode: pool1 112->56, pool2 56->28; padded 3-D {6,5,4} k3 s2 p1 -> {4,3,3}
    # (test_pooling_layer.cpp:1414-1431) and unpadded -> {3,2,2} (:1397-1412)
    assert refnet.pool_out_shape([112, 112], [3, 3], [2, 2], [0, 0]) == [56, 56]
    assert refnet.pool_out_shape([56, 56], [3, 3], [2, 2], [0, 0]) == [28, 28]
    assert refnet.pool_out_shape([6, 5, 4], [3, 3, 3], [2, 2, 2], [1, 1, 1]) == [4, 3, 3]
    assert refnet.pool_out_shape([6, 5, 4], [3, 3, 3], [2, 2, 2], [0, 0, 0]) == [3, 2, 2]
    assert refnet.pool_out_shape([28, 28], [3, 3], [1, 1], [1, 1]) == [28, 28]


def test_bn_frozen_closed_form():
    # test_bn_layer.cpp:107-154 (TestForwardFrozen): running mean = c, running var = c,
    # slope 1, bias 0 -> y = (x - c) / sqrt(c + eps); the reference checks against eps-free form at 1e-3... we use eps.
    rng = np.random.default_rng(1701)
    x = rng.normal(size=(5, 2, 3, 4)).astype(np.float32)
    for c in (0.5, 2.0):
        ch = x.shape[1]
        y = refnet.bn_forward_test(x, np.ones(ch), np.zeros(ch), np.full(ch, c), np.full(ch, c), 1e-5)
        assert np.allclose(y, (x - c) / np.sqrt(c + 1e-5), atol=1e-5)


def test_bn_train_zero_mean_unit_var():
    # test_bn_layer.cpp:44-87 (TestForward): per-channel output mean ~ 0, var ~ 1 (tol 1e-3)
    rng = np.random.default_rng(1701)
    x = rng.normal(2.0, 3.0, size=(5, 2, 3, 4)).astype(np.float32)
    rm, rv = np.zeros(2, np.float32), np.zeros(2, np.float32)
    y, bm, bv = refnet.bn_forward_train(x, np.ones(2), np.zeros(2), rm, rv)
    for c in range(2):
        assert abs(y[:, c].mean()) < 1e-3
        assert abs(y[:, c].var() - 1.0) < 1e-3
    # running <- (1 - 0.9) * batch + 0.9 * running   (bn_layer.cpp:118-122,153-156)
    assert np.allclose(rm, 0.1 * bm) and np.allclose(rv, 0.1 * bv)
    # 5-D input: same statistics over N*D*H*W
    x5 = x.reshape(5, 2, 3, 2, 2)
    y5, _, _ = refnet.bn_forward_train(x5, np.ones(2), np.zeros(2), np.zeros(2, np.float32), np.zeros(2, np.float32))
    assert np.allclose(y5.reshape(y.shape), y, atol=1e-6)


@pytest.mark.parametrize("shape,k,s,p", [
    ((2, 3, 6, 4), [3, 3], [2, 2], [0, 0]),         # test_convolution_layer.cpp TestSimpleConvolution geometry
    ((2, 3, 5, 5, 5), [3, 3, 3], [2, 2, 2], [0, 0, 0]),  # :300-345 TestSimple3DConvolution
    ((1, 4, 7, 9), [1, 1], [1, 1], [0, 0]),         # :348 Test1x1Convolution
    ((1, 2, 4, 6, 5), [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ((2, 3, 9, 9), [7, 7], [2, 2], [3, 3]),
])
def test_conv_im2col_path_matches_naive_spec(shape, k, s, p):
    # the reference checks ConvolutionLayer against its naive caffe_conv (test_convolution_layer.cpp:18-134)
    # at 1e-4; the oracle carries both restatements and they must agree.
    rng = np.random.default_rng(7)
    x = rng.normal(size=shape).astype(np.float32)
    w = rng.normal(size=[4, shape[1]] + k).astype(np.float32)
    b = rng.normal(size=4).astype(np.float32)
    y1 = refnet.conv_forward(x, w, b, k, s, p)
    y2 = refnet.conv_forward(x, w, b, k, s, p, naive=True)
    assert y1.shape == y2.shape
    assert np.abs(y1 - y2).max() < 1e-4


def test_conv_sobel_separable():
    # test_convolution_layer.cpp:403-494 (TestSobelConvolution): a 3x3 Sobel filter equals the
    # composition of a 3x1 smoothing [1 2 1]^T and a 1x3 difference [-1 0 1].
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 1, 8, 9)).astype(np.float32)
    sob = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32).reshape(1, 1, 3, 3)
    full = refnet.conv_forward(x, sob, None, [3, 3], [1, 1], [0, 0])
    col = np.array([1, 2, 1], np.float32).reshape(1, 1, 3, 1)
    row = np.array([-1, 0, 1], np.float32).reshape(1, 1, 1, 3)
    sep = refnet.conv_forward(refnet.conv_forward(x, col, None, [3, 1], [1, 1], [0, 0]), row, None, [1, 3], [1, 1], [0, 0])
    assert np.abs(full - sep).max() < 1e-4


def test_conv_nd_equals_2d():
    # test_convolution_layer.cpp:496-613 (TestNDAgainst2D): a 3-D conv with depth-1 kernel on a depth-1 volume == 2-D conv
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 3, 6, 7)).astype(np.float32)
    w = rng.normal(size=(5, 3, 3, 3)).astype(np.float32)
    b = rng.normal(size=5).astype(np.float32)
    y2 = refnet.conv_forward(x, w, b, [3, 3], [1, 1], [1, 1])
    y3 = refnet.conv_forward(x[:, :, None], w[:, :, None], b, [1, 3, 3], [1, 1, 1], [0, 1, 1])
    assert np.abs(y3[:, :, 0] - y2).max() < 1e-5


def test_eltwise_concat_permute_ip():
    rng = np.random.default_rng(11)
    a = rng.normal(size=(2, 3, 4, 5)).astype(np.float32)
    b = rng.normal(size=(2, 3, 4, 5)).astype(np.float32)
    assert np.allclose(refnet.eltwise_sum(a, b), a + b)                    # test_eltwise_layer.cpp:87-104
    x = rng.normal(size=(2, 4, 3, 5, 6)).astype(np.float32)
    assert np.array_equal(refnet.permute(x, [0, 2, 1, 3, 4]), x.transpose(0, 2, 1, 3, 4))  # permute_layer.cpp:9-26
    w = rng.uniform(size=(10, 60)).astype(np.float32)
    bias = rng.uniform(1, 2, size=10).astype(np.float32)
    y = refnet.inner_product(np.abs(a), w, bias)
    assert y.shape == (2, 10) and (y >= 1).all()                           # test_inner_product_layer.cpp:58-86
    assert np.allclose(y, np.abs(a).reshape(2, -1) @ w.T + bias, atol=1e-4)


def test_bn_fold_identity():
    # caffe_3d/python/gen_bn_inference.py:121-134 fold, :23-33 allclose check:
    # BN(conv(x; W, b)) == conv(x; W*g*invstd, (b-mu)*g*invstd + beta)
    rng = np.random.default_rng(13)
    x = rng.normal(size=(2, 3, 6, 6)).astype(np.float32)
    w = rng.normal(size=(4, 3, 3, 3)).astype(np.float32)
    b = rng.normal(size=4).astype(np.float32)
    g, beta = rng.uniform(0.5, 1.5, 4).astype(np.float32), rng.normal(size=4).astype(np.float32)
    mu, var = rng.normal(size=4).astype(np.float32), rng.uniform(0.5, 1.5, 4).astype(np.float32)
    y = refnet.bn_forward_test(refnet.conv_forward(x, w, b, [3, 3], [1, 1], [1, 1]), g, beta, mu, var, 1e-5)
    s = g / np.sqrt(var + 1e-5)
    yf = refnet.conv_forward(x, w * s[:, None, None, None], (b - mu) * s + beta, [3, 3], [1, 1], [1, 1])
    assert np.allclose(y, yf, atol=1e-4)
