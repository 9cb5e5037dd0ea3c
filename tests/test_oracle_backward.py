"""The oracle's backward pass (oracle/refnet.py RefNet.backward over ref_cpu.c ref_conv_backward / ref_pool_backward /
ref_bn_backward_train) checked the way the reference checks its layers: central finite differences of the loss against
the analytic gradient (caffe_3d/include/caffe/test/test_gradient_check_util.hpp:18-60, step 1e-2 / threshold 1e-3 there;
the whole-net variant here perturbs parameters and inputs of a small TRAIN-phase net), plus closed forms."""
import numpy as np
import pytest

from oracle import refnet

NET = '''name: "t"
input: "data" input_dim: 2 input_dim: 3 input_dim: 9 input_dim: 9
input: "label" input_dim: 2 input_dim: 1 input_dim: 1 input_dim: 1
layer { name: "c1" type: "Convolution" bottom: "data" top: "c1" convolution_param { num_output: 8 kernel_size: 3 stride: 2 pad: 1 } }
layer { name: "c1_bn" type: "BN" bottom: "c1" top: "c1_bn" }
layer { name: "c1_relu" type: "ReLU" bottom: "c1_bn" top: "c1_bn" }
layer { name: "p" type: "Pooling" bottom: "c1_bn" top: "p" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }
layer { name: "c2" type: "Convolution" bottom: "p" top: "c2" convolution_param { num_output: 8 kernel_size: 1 } }
layer { name: "c3" type: "Convolution" bottom: "p" top: "c3" convolution_param { num_output: 8 kernel_size: 3 pad: 1 } }
layer { name: "s" type: "Eltwise" bottom: "c2" bottom: "c3" top: "s" }
layer { name: "s_bn" type: "BN" bottom: "s" top: "s_bn" }
layer { name: "s_relu" type: "ReLU" bottom: "s_bn" top: "s_bn" }
layer { name: "a" type: "Pooling" bottom: "s_bn" top: "a" pooling_param { pool: AVE kernel_size: 3 stride: 1 pad: 1 } }
layer { name: "cat" type: "Concat" bottom: "a" bottom: "s_bn" top: "cat" }
layer { name: "gp" type: "Pooling" bottom: "cat" top: "gp" pooling_param { pool: AVE kernel_size: 2 stride: 1 } }
layer { name: "drop" type: "Dropout" bottom: "gp" top: "gp" dropout_param { dropout_ratio: 0.3 } }
layer { name: "fc" type: "InnerProduct" bottom: "gp" top: "fc" inner_product_param { num_output: 5 } }
layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc" bottom: "label" top: "loss" }
'''


def _setup():
    net = refnet.RefNet(NET, phase="TRAIN").init_params(3)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 3, 9, 9)).astype(np.float32)
    lab = np.array([1, 3], np.float32).reshape(2, 1, 1, 1)
    mask = (rng.random((2, 16, 1, 1)) < 0.7).astype(np.float32)
    return net, x, lab, {"drop": mask}


def _loss(params, x, lab, masks):
    n = refnet.RefNet(NET, phase="TRAIN").set_params(params)
    return float(n.forward({"data": x, "label": lab}, dropout_masks=masks)["loss"])


def test_whole_net_gradients_match_finite_differences():
    net, x, lab, masks = _setup()
    P0 = net.params_dict()
    net.forward({"data": x, "label": lab}, dropout_masks=masks)
    net.set_params(P0)  # the forward pass moved the BN running averages (they do not enter the TRAIN loss)
    _, pd = net.backward()
    rng = np.random.default_rng(1)
    eps = 2e-3
    checked = 0
    for lname in ("c1", "c2", "c3", "c1_bn", "s_bn", "fc"):
        for bi in range(2):
            arr = P0[lname][bi]
            for _ in range(4):
                idx = tuple(int(rng.integers(0, s)) for s in arr.shape)
                Pp = {k: [a.copy() for a in v] for k, v in P0.items()}
                Pm = {k: [a.copy() for a in v] for k, v in P0.items()}
                Pp[lname][bi][idx] += eps
                Pm[lname][bi][idx] -= eps
                fd = (_loss(Pp, x, lab, masks) - _loss(Pm, x, lab, masks)) / (2 * eps)
                an = float(pd[lname][bi][idx])
                assert abs(fd - an) <= 2e-3 * max(1.0, abs(fd), abs(an)) + 2e-4, (lname, bi, idx, fd, an)
                checked += 1
    assert checked == 48
    # a bias in front of a batch-statistics BN cannot change the loss (the mean is subtracted): its gradient is ~0
    assert np.abs(pd["c1"][1]).max() < 1e-5 and np.abs(pd["c3"][1]).max() < 1e-5
    # BN running statistics carry no gradient (lr_mult forced to 0, bn_layer.cpp:46-53)
    assert np.abs(pd["s_bn"][2]).max() == 0 and np.abs(pd["s_bn"][3]).max() == 0


def test_conv_backward_is_the_adjoint_of_forward():
    # <conv(x, w), dy> == <x, dgrad(dy)> == <w, wgrad(x, dy)>  (linearity; 2-D stride 2 and 3-D)
    rng = np.random.default_rng(5)
    for shape, cout, k, s, p in (((2, 5, 9, 8), 7, [3, 3], [2, 2], [1, 1]), ((1, 4, 4, 6, 5), 6, [3, 3, 3], [1, 2, 1], [1, 0, 1])):
        x = rng.normal(size=shape).astype(np.float32)
        w = rng.normal(size=[cout, shape[1]] + k).astype(np.float32)
        y = refnet.conv_forward(x, w, None, k, s, p)
        dy = rng.normal(size=y.shape).astype(np.float32)
        dx, dw, db = refnet.conv_backward(x, w, dy, k, s, p)
        lhs = float((y.astype(np.float64) * dy).sum())
        assert abs(lhs - float((x.astype(np.float64) * dx).sum())) <= 1e-4 * abs(lhs)
        assert abs(lhs - float((w.astype(np.float64) * dw).sum())) <= 1e-4 * abs(lhs)
        assert np.allclose(db, dy.sum(axis=tuple(i for i in range(dy.ndim) if i != 1)), rtol=1e-4, atol=1e-4)


def test_max_pool_backward_routes_to_the_first_maximum():
    # pooling_layer.cpp:206-222 records the first maximum in scan order; ties must not split or move the gradient
    x = np.zeros((1, 1, 4, 4), np.float32)
    x[0, 0, 1, 1] = 1.0
    x[0, 0, 1, 2] = 1.0  # tie inside the first window
    dy = np.arange(1, 5, dtype=np.float32).reshape(1, 1, 2, 2)
    dx = refnet.pool_backward(x, dy, [3, 3], [2, 2], [0, 0], "MAX")
    # windows: (0,0) rows0-2 cols0-2 -> first max (1,1); (0,1) cols 2-3 (clipped): max (1,2); (1,0) rows2-3: all zero -> (2,0); (1,1): (2,2)
    want = np.zeros_like(x)
    want[0, 0, 1, 1] = 1
    want[0, 0, 1, 2] = 2
    want[0, 0, 2, 0] = 3
    want[0, 0, 2, 2] = 4
    assert np.array_equal(dx, want)


def test_ave_pool_backward_uses_the_pad_inclusive_divisor():
    # pooling_layer.cpp:340-356: every window of a 3x3/s1/p1 pool divides by 9, also at the border
    x = np.zeros((1, 1, 3, 3), np.float32)
    dy = np.ones((1, 1, 3, 3), np.float32)
    dx = refnet.pool_backward(x, dy, [3, 3], [1, 1], [1, 1], "AVE")
    assert np.allclose(dx[0, 0], np.array([[4, 6, 4], [6, 9, 6], [4, 6, 4]], np.float32) / 9.0)


def test_bn_train_backward_properties():
    # the gradient w.r.t. x of a batch-normalised tensor is orthogonal to 1 and to x_norm per channel
    rng = np.random.default_rng(2)
    x = rng.normal(size=(3, 4, 5, 5)).astype(np.float32)
    slope = rng.uniform(0.5, 1.5, 4).astype(np.float32)
    rm, rv = np.zeros(4, np.float32), np.zeros(4, np.float32)
    _, bm, bv = refnet.bn_forward_train(x, slope, np.zeros(4, np.float32), rm, rv, 0.9, 1e-5)
    dy = rng.normal(size=x.shape).astype(np.float32)
    dx, ds, db = refnet.bn_backward_train(x, dy, slope, bm, bv, 1e-5)
    xn = (x - bm.reshape(1, 4, 1, 1)) / np.sqrt(bv.reshape(1, 4, 1, 1) + 1e-5)
    assert np.abs(dx.sum((0, 2, 3))).max() < 1e-3
    assert np.abs((dx * xn).sum((0, 2, 3))).max() < 1e-3
    assert np.allclose(ds, (dy * xn).sum((0, 2, 3)), rtol=1e-4, atol=1e-4)
    assert np.allclose(db, dy.sum((0, 2, 3)), rtol=1e-4, atol=1e-4)
    # running averages: (1 - m) * batch + m * running  (bn_layer.cpp:120-123)
    assert np.allclose(rm, 0.1 * bm, atol=1e-6) and np.allclose(rv, 0.1 * bv, atol=1e-6)
