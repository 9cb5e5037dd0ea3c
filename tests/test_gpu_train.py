"""GPU parity tests of the training path: TRAIN-phase forward
(batch-statistics BN, Dropout, SoftmaxWithLoss, Accuracy) and Backward through the public surface
(caffe shim -> C ABI -> sm_100a kernels: tcgen05 dgrad / wgrad GEMMs + the HBM-bound backward kernels) against the
oracle's backward pass (oracle/refnet.py, finite-difference-checked in tests/test_oracle_backward.py).

Tolerances.  The device stores activations and feature-map gradients as bf16 and accumulates in fp32; parameters and
their gradients are fp32.
  * single ops are fed operands that are exactly representable in bf16 (the oracle sees the SAME numbers), so the only
    device error is the output rounding: bf16 blobs within 1 bf16 ulp (check_bf16_blob), fp32 parameter gradients
    within TOL_WGRAD of the largest entry (summation order only);
  * whole nets are checked teacher-forced in both directions (each layer's forward from the device's own bottom blobs,
    each layer's backward from the device's own top gradient): blob data / gradients rel-L2 <= TOL_STEP_BLOB (one bf16
    rounding), parameter gradients <= TOL_STEP_PARAM.  The free-running deviation from the fp32 oracle end to end is
    printed next to it: with random weights, batch-statistics BN over a two-clip batch and ~1 % ReLU-mask flips per
    layer it grows to 0.1 (activations) / 0.7 (gradients of the first layers) -- chaotic amplification of bf16 storage
    noise in this harness, not an arithmetic error of any op (every step is accurate to 2e-3)."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import check_bf16_blob, describe_mismatch, load_params, rel_l2, rel_max

pytestmark = pytest.mark.gpu

TOL_WGRAD = 2e-4       # fp32 parameter gradients from identical bf16 operands: accumulation order only


def train_net(txt, **opts):
    import caffe
    return caffe.Net.from_string(txt, caffe.TRAIN, **opts)


def bf(a):
    return refnet.round_bf16(np.asarray(a, np.float32))


def lst(v):
    return "[%s]" % ", ".join(str(i) for i in v)


def header(shape, label=False):
    txt = 'name: "t"\ninput: "data"\ninput_shape { %s }\n' % " ".join("dim: %d" % d for d in shape)
    if label:
        txt += 'input: "label"\ninput_shape { dim: %d dim: 1 dim: 1 dim: 1 }\n' % shape[0]
    return txt


def conv(name, bottom, cout, k, s, p, top=None):
    return ('layer { name: "%s" type: "Convolution" bottom: "%s" top: "%s" convolution_param { num_output: %d '
            'kernel_size: %s stride: %s pad: %s } }\n' % (name, bottom, top or name, cout, lst(k), lst(s), lst(p)))


# ---------------------------------------------------------------------------------------------------------------
CONV_BWD = [
    # shape of data, cmid (channels of the conv input), cout, k, s, p
    ("2d_3x3", (2, 8, 12, 12), 64, 64, [3, 3], [1, 1], [1, 1]),
    ("2d_1x1_cin192", (2, 8, 14, 14), 192, 64, [1, 1], [1, 1], [0, 0]),
    ("2d_3x3_s2_odd", (2, 8, 15, 15), 64, 160, [3, 3], [2, 2], [1, 1]),
    ("2d_cin96_cout96", (2, 8, 14, 14), 96, 96, [3, 3], [1, 1], [1, 1]),
    ("2d_cout32", (3, 8, 10, 10), 64, 32, [1, 1], [1, 1], [0, 0]),
    ("3d_3x3x3", (1, 8, 4, 10, 10), 64, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ("3d_s2_res4a", (2, 8, 8, 14, 14), 128, 256, [3, 3, 3], [2, 2, 2], [1, 1, 1]),
    ("3d_cin256_cout512", (2, 8, 2, 7, 7), 256, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ("3d_1x1_s2_down", (2, 8, 4, 8, 8), 128, 256, [1, 1, 1], [2, 2, 2], [0, 0, 0]),
    ("3d_s2_tiny_m64", (2, 8, 2, 4, 4), 32, 64, [3, 3, 3], [2, 2, 2], [1, 1, 1]),   # fewer positions than one 128-row tile
    ("2d_1x1_s2_compact", (2, 8, 15, 12), 64, 96, [1, 1], [2, 2], [0, 0]),   # strided 1x1: dgrad on the compact grid + scatter
    ("3d_s1_tiny_m64", (2, 8, 2, 4, 4), 48, 32, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
]


@pytest.mark.parametrize("case", CONV_BWD, ids=[c[0] for c in CONV_BWD])
def test_conv_backward(gpu, case):
    """dgrad + wgrad + bias gradient of one convolution, operands exactly representable in bf16."""
    _, shape, cmid, cout, k, s, p = case
    nsp = len(shape) - 2
    txt = header(shape) + conv("a", "data", cmid, [1] * nsp, [1] * nsp, [0] * nsp) + conv("c", "a", cout, k, s, p)
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(3)
    rng = np.random.default_rng(0)
    x = bf(rng.normal(size=shape))
    P = ref.params_dict()
    for name in P:
        P[name] = [bf(P[name][0]), P[name][1]]  # weights exactly representable; the bias is added in fp32 on both sides
    ref.set_params(P)
    net = train_net(txt)
    load_params(net, P)
    net.blobs["data"].data[...] = x
    net.forward()
    a_dev = net.blobs["a"].data.copy()  # what the convolution under test actually read (bf16 values)
    c_dev = net.blobs["c"].data.copy()
    dy = bf(rng.normal(size=c_dev.shape))
    net.clear_param_diffs()
    net.backward(**{"c": dy})
    dx_want, dw_want, db_want = refnet.conv_backward(a_dev, P["c"][0], dy, k, s, p)
    dw = net.params["c"][0].diff.copy()
    db = net.params["c"][1].diff.copy()
    assert rel_max(dw, dw_want) <= TOL_WGRAD, describe_mismatch(dw, dw_want, "dW")
    assert rel_max(db, db_want) <= TOL_WGRAD, describe_mismatch(db, db_want, "db")
    check_bf16_blob(net.blobs["a"].diff.copy(), dx_want, "dX")
    # caffe accumulates parameter gradients over backward calls (beta = 1) until they are cleared
    net.backward(**{"c": dy})
    assert rel_max(net.params["c"][0].diff, 2 * dw_want) <= TOL_WGRAD
    # the 1x1 layer in front: its gradients come from the device's own dX
    dxa = bf(net.blobs["a"].diff.copy())
    _, dwa_want, dba_want = refnet.conv_backward(x, P["a"][0], dxa, [1] * nsp, [1] * nsp, [0] * nsp, need_dx=False)
    net.clear_param_diffs()
    net.backward(**{"c": dy})
    assert rel_max(net.params["a"][0].diff, dwa_want) <= TOL_WGRAD, describe_mismatch(net.params["a"][0].diff, dwa_want, "dW(a)")


def test_two_consumers_accumulate_into_one_gradient(gpu):
    """a blob read by two layers: the auto-inserted Split sums their gradients (split_layer.cpp); here the second
    dgrad adds into the buffer the first one wrote (residual operand of the GEMM epilogue)"""
    shape = (2, 8, 2, 8, 8)
    one, zero = [1, 1, 1], [0, 0, 0]
    two = [2, 2, 2]
    # (g, f: strided 1x1 convolutions -- their dgrad runs on the compact grid and is scattered; g adds into the buffer, f,
    # the last layer = the first one to run backward, overwrites it)
    txt = header(shape) + conv("a", "data", 32, one, one, zero) + conv("c", "a", 64, [3, 3, 3], [2, 2, 2], [1, 1, 1]) + \
        conv("g", "a", 32, one, two, zero) + conv("d", "a", 48, [3, 3, 3], [2, 2, 2], [1, 1, 1]) + conv("e", "a", 16, one, one, zero) + \
        conv("f", "a", 24, one, two, zero)
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(8)
    P = {k: [bf(v[0]), v[1]] for k, v in ref.params_dict().items()}
    rng = np.random.default_rng(2)
    x = bf(rng.normal(size=shape))
    net = train_net(txt)
    load_params(net, P)
    net.blobs["data"].data[...] = x
    net.forward()
    a_dev = net.blobs["a"].data.copy()
    dys = {n: bf(rng.normal(size=net.blobs[n].data.shape)) for n in ("c", "g", "d", "e", "f")}
    net.clear_param_diffs()
    net.backward(**dys)
    total = np.zeros_like(a_dev)
    parts = {}
    for n, (k, s_, p_) in {"c": ([3, 3, 3], [2, 2, 2], [1, 1, 1]), "d": ([3, 3, 3], [2, 2, 2], [1, 1, 1]), "e": (one, one, zero),
                           "g": (one, two, zero), "f": (one, two, zero)}.items():
        dx, dw, _ = refnet.conv_backward(a_dev, P[n][0], dys[n], k, s_, p_)
        parts[n] = dx
        total += dx
        assert rel_max(net.params[n][0].diff, dw) <= TOL_WGRAD, describe_mismatch(net.params[n][0].diff, dw, "dW " + n)
    got = net.blobs["a"].diff.copy()
    # three bf16 roundings (one per accumulation step) instead of one: within 3 ulp-ish -> compare in rel-L2
    assert rel_l2(got, total) <= 8e-3, describe_mismatch(got, total, "d(a) = sum of three consumers; vs single parts: %s" % {
        n: rel_l2(got, v) for n, v in parts.items()})


def test_stem_conv_backward(gpu):
    """conv1 7x7/s2/p3 over 3 channels: its weight gradient runs over the space-to-depth cells of the forward pass"""
    shape = (4, 3, 64, 64)
    txt = header(shape) + conv("conv1", "data", 64, [7, 7], [2, 2], [3, 3])
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(5)
    P = {k: [bf(v[0]), v[1]] for k, v in ref.params_dict().items()}
    rng = np.random.default_rng(1)
    x = bf(rng.normal(size=shape))
    net = train_net(txt)
    load_params(net, P)
    net.blobs["data"].data[...] = x
    net.forward()
    y = net.blobs["conv1"].data.copy()
    want = refnet.conv_forward(x, P["conv1"][0], P["conv1"][1], [7, 7], [2, 2], [3, 3])
    check_bf16_blob(y, want, "conv1")
    dy = bf(rng.normal(size=y.shape))
    net.clear_param_diffs()
    net.backward(**{"conv1": dy})
    _, dw_want, db_want = refnet.conv_backward(x, P["conv1"][0], dy, [7, 7], [2, 2], [3, 3], need_dx=False)
    dw = net.params["conv1"][0].diff.copy()
    assert rel_max(dw, dw_want) <= TOL_WGRAD, describe_mismatch(dw, dw_want, "dW")
    assert rel_max(net.params["conv1"][1].diff, db_want) <= TOL_WGRAD


@pytest.mark.parametrize("shape", [(4, 8, 9, 9), (2, 8, 4, 6, 6)], ids=["2d", "3d"])
def test_bn_train_forward_backward(gpu, shape):
    nsp = len(shape) - 2
    txt = header(shape) + conv("c", "data", 32, [1] * nsp, [1] * nsp, [0] * nsp)
    txt += 'layer { name: "c_bn" type: "BN" bottom: "c" top: "c_bn" bn_param { momentum: 0.8 } }\n'
    txt += 'layer { name: "c_relu" type: "ReLU" bottom: "c_bn" top: "c_bn" }\n'
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(2)
    P = ref.params_dict()
    rng = np.random.default_rng(3)
    x = bf(rng.normal(size=shape) * 2 + 0.5)
    net = train_net(txt)
    load_params(net, P)
    net.blobs["data"].data[...] = x
    net.forward()
    c_dev = net.blobs["c"].data.copy()      # the BN layer's input as the device stored it
    y_dev = net.blobs["c_bn"].data.copy()
    rm, rv = P["c_bn"][2].ravel().copy(), P["c_bn"][3].ravel().copy()
    y_want, bm, bv = refnet.bn_forward_train(c_dev, P["c_bn"][0], P["c_bn"][1], rm, rv, 0.8, 1e-5)
    check_bf16_blob(y_dev, np.maximum(y_want, 0), "c_bn")
    # running statistics: (1 - m) * batch + m * running   (bn_layer.cpp:120-123,151-153)
    assert rel_max(net.params["c_bn"][2].data.ravel(), rm) <= 1e-4
    assert rel_max(net.params["c_bn"][3].data.ravel(), rv) <= 1e-4
    dy = bf(rng.normal(size=y_dev.shape))
    net.clear_param_diffs()
    net.backward(**{"c_bn": dy})
    dy_eff = dy * (y_dev > 0)
    dx_want, ds_want, dbias_want = refnet.bn_backward_train(c_dev, dy_eff, P["c_bn"][0], bm, bv, 1e-5)
    assert rel_max(net.params["c_bn"][0].diff.ravel(), ds_want) <= 1e-3
    assert rel_max(net.params["c_bn"][1].diff.ravel(), dbias_want) <= 1e-3
    check_bf16_blob(net.blobs["c"].diff.copy(), dx_want, "d(c)")
    assert np.abs(net.params["c_bn"][2].diff).max() == 0 and np.abs(net.params["c_bn"][3].diff).max() == 0


POOL_BWD = [
    ("max_3x3_s2_ceil", (2, 16, 13, 13), "MAX", [3, 3], [2, 2], [0, 0]),
    ("ave_3x3_s1_p1", (2, 16, 9, 9), "AVE", [3, 3], [1, 1], [1, 1]),
    ("max_3x3_s1_p1", (1, 8, 7, 7), "MAX", [3, 3], [1, 1], [1, 1]),
    ("ave_3d", (2, 8, 4, 6, 6), "AVE", [2, 3, 3], [2, 1, 1], [0, 1, 1]),
]


@pytest.mark.parametrize("case", POOL_BWD, ids=[c[0] for c in POOL_BWD])
def test_pool_backward(gpu, case):
    _, shape, method, k, s, p = case
    nsp = len(shape) - 2
    txt = header(shape) + conv("c", "data", shape[1], [1] * nsp, [1] * nsp, [0] * nsp)
    txt += ('layer { name: "p" type: "Pooling" bottom: "c" top: "p" pooling_param { pool: %s kernel_size: %s stride: %s '
            'pad: %s } }\n' % (method, lst(k), lst(s), lst(p)))
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(4)
    rng = np.random.default_rng(6)
    # coarse values: many exact ties, so the first-maximum rule of the reference's max_idx_ matters
    x = np.round(rng.normal(size=shape) * 2).astype(np.float32)
    P = ref.params_dict()
    P["c"][0] = np.eye(shape[1], dtype=np.float32).reshape(P["c"][0].shape)
    P["c"][1] = np.zeros_like(P["c"][1])
    net = train_net(txt)
    load_params(net, P)
    net.blobs["data"].data[...] = x
    net.forward()
    c_dev = net.blobs["c"].data.copy()
    assert np.array_equal(c_dev, x)
    y = net.blobs["p"].data.copy()
    dy = bf(rng.normal(size=y.shape))
    net.backward(**{"p": dy})
    want = refnet.pool_backward(c_dev, dy, k, s, p, method)
    check_bf16_blob(net.blobs["c"].diff.copy(), want, "d(c)")


# ---------------------------------------------------------------------------------------------------------------
TOY = header((4, 16, 33, 33)) + 'input: "label"\ninput_shape { dim: 2 dim: 1 dim: 1 dim: 1 }\n' + \
    conv("c1", "data", 32, [3, 3], [2, 2], [1, 1]) + \
    'layer { name: "c1_bn" type: "BN" bottom: "c1" top: "c1_bn" }\n' \
    'layer { name: "c1_relu" type: "ReLU" bottom: "c1_bn" top: "c1_bn" }\n' \
    'layer { name: "p" type: "Pooling" bottom: "c1_bn" top: "p" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }\n' + \
    conv("b1", "p", 16, [1, 1], [1, 1], [0, 0]) + \
    'layer { name: "b1_bn" type: "BN" bottom: "b1" top: "b1_bn" }\n' \
    'layer { name: "b1_relu" type: "ReLU" bottom: "b1_bn" top: "b1_bn" }\n' + \
    conv("b2r", "p", 16, [1, 1], [1, 1], [0, 0]) + \
    'layer { name: "b2r_bn" type: "BN" bottom: "b2r" top: "b2r_bn" }\n' \
    'layer { name: "b2r_relu" type: "ReLU" bottom: "b2r_bn" top: "b2r_bn" }\n' + \
    conv("b2", "b2r_bn", 24, [3, 3], [1, 1], [1, 1]) + \
    'layer { name: "b2_bn" type: "BN" bottom: "b2" top: "b2_bn" }\n' \
    'layer { name: "b2_relu" type: "ReLU" bottom: "b2_bn" top: "b2_bn" }\n' \
    'layer { name: "pa" type: "Pooling" bottom: "p" top: "pa" pooling_param { pool: AVE kernel_size: 3 stride: 1 pad: 1 } }\n' + \
    conv("b3", "pa", 8, [1, 1], [1, 1], [0, 0]) + \
    'layer { name: "b3_bn" type: "BN" bottom: "b3" top: "b3_bn" }\n' \
    'layer { name: "b3_relu" type: "ReLU" bottom: "b3_bn" top: "b3_bn" }\n' \
    'layer { name: "cat" type: "Concat" bottom: "b1_bn" bottom: "b2_bn" bottom: "b3_bn" top: "cat" }\n' \
    'layer { name: "r3" type: "Reshape" bottom: "cat" top: "v" reshape_param { shape { dim: -1 dim: 2 dim: 48 dim: 8 dim: 8 } } }\n' \
    'layer { name: "tr" type: "Permute" bottom: "v" top: "vt" permute_param { order: [0,2,1,3,4] } }\n' + \
    conv("ra", "vt", 32, [3, 3, 3], [1, 1, 1], [1, 1, 1]) + \
    'layer { name: "ra_bn" type: "BN" bottom: "ra" top: "ra_bn" }\n' \
    'layer { name: "ra_relu" type: "ReLU" bottom: "ra_bn" top: "ra_bn" }\n' + \
    conv("rb", "ra_bn", 32, [3, 3, 3], [1, 1, 1], [1, 1, 1]) + \
    'layer { name: "res" type: "Eltwise" bottom: "rb" bottom: "ra" top: "res" }\n' \
    'layer { name: "res_bn" type: "BN" bottom: "res" top: "res_bn" }\n' \
    'layer { name: "res_relu" type: "ReLU" bottom: "res_bn" top: "res_bn" }\n' + \
    conv("rc", "res_bn", 64, [3, 3, 3], [2, 2, 2], [1, 1, 1]) + \
    conv("rd", "res_bn", 64, [3, 3, 3], [2, 2, 2], [1, 1, 1]) + \
    'layer { name: "res2" type: "Eltwise" bottom: "rc" bottom: "rd" top: "res2" }\n' \
    'layer { name: "res2_bn" type: "BN" bottom: "res2" top: "res2_bn" }\n' \
    'layer { name: "res2_relu" type: "ReLU" bottom: "res2_bn" top: "res2_bn" }\n' \
    'layer { name: "gp" type: "Pooling" bottom: "res2_bn" top: "gp" pooling_param { pool: AVE kernel_size: [1, 4, 4] } }\n' \
    'layer { name: "gp_r" type: "Reshape" bottom: "gp" top: "gp_r" reshape_param { shape { dim: -1 dim: 64 } } }\n' \
    'layer { name: "drop" type: "Dropout" bottom: "gp_r" top: "gp_r" dropout_param { dropout_ratio: 0.25 } }\n' \
    'layer { name: "fc" type: "InnerProduct" bottom: "gp_r" top: "fc" inner_product_param { num_output: 10 } }\n' \
    'layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc" bottom: "label" top: "loss" }\n'


TOL_STEP_BLOB = 8e-3   # teacher-forced: one backward step from the device's own inputs -> bf16 output rounding (+ accumulation)
TOL_STEP_PARAM = 3e-3  # teacher-forced fp32 parameter gradients


def run_train_net_against_oracle(txt, inputs, seed, dropout_layer, dropout_blob, skip_data=(), skip_diff=()):
    """Forward + backward on the device, then the oracle teacher-forced in BOTH directions: every layer's forward is
    recomputed from the device's bottom blobs and every layer's backward from the device's top gradient, so each op is
    judged on the inputs it actually had (no chaotic error growth through ReLU-mask flips and small-batch BN).
    Free-running numbers (oracle fp32 end to end) are printed for the record."""
    ref = refnet.RefNet(txt, phase="TRAIN").init_params(seed)
    P = ref.params_dict()
    net = train_net(txt)
    load_params(net, P)
    for k, v in inputs.items():
        net.blobs[k].data[...] = v
    loss = float(net.forward()["loss"])
    net.clear_param_diffs()
    net.backward()
    names = [b for b in net._blob_names if "_split_" not in b]
    dev, ddev = {}, {}
    for b in names:
        try:
            dev[b] = net.blobs[b].data.copy()
        except RuntimeError:
            continue
        if b not in inputs:
            try:
                ddev[b] = net.blobs[b].diff.copy()
            except RuntimeError:
                pass
    y = dev[dropout_blob]
    masks = {dropout_layer: (y != 0).astype(np.float32)}
    # free-running reference (fp32 end to end)
    ref.set_params(P)
    want = ref.forward(inputs, dropout_masks=masks)
    free_d, free_p = ref.backward()
    assert abs(loss - float(want["loss"])) <= 2e-2 * max(1.0, abs(float(want["loss"]))), (loss, float(want["loss"]))
    # teacher-forced reference
    ref.set_params(P)
    # (views of the inputs and the blob an in-place Dropout rewrites are not fed back: the device keeps caffe's in-place
    # semantics there -- data AND diff of that blob are the rewritten values -- which the substitution would apply twice)
    teach = {k: v for k, v in dev.items() if k not in inputs and k != "loss" and k not in skip_data}
    # (skip_diff: a Reshape view of a blob with a second consumer -- caffe puts a Split in front, so the view's own diff
    # holds one branch only, while the device keeps ONE diff buffer for the storage and sums both branches into it; the
    # sum is checked on the viewed blob, the view itself is neither compared nor fed back)
    ddev = {k: v for k, v in ddev.items() if k not in skip_data and k not in skip_diff}
    forced = ref.forward(inputs, dropout_masks=masks, teacher=teach)
    bn_after = {l.name: [p.copy() for p in l.params] for l in ref.layers if l.type == "BN"}
    own_d, own_p = ref.backward(teacher_diffs=ddev)
    report, failures = [], []
    for b, g in ddev.items():
        if b not in own_d or b == "loss":
            continue
        w = own_d[b].reshape(g.shape)
        if np.abs(w).max() == 0 and np.abs(g).max() == 0:
            continue
        e = rel_l2(g, w)
        fr = rel_l2(g, free_d[b].reshape(g.shape)) if b in free_d else float("nan")
        report.append(("d(%s)" % b, e, fr))
        if not e <= TOL_STEP_BLOB:
            failures.append(describe_mismatch(g, w, "d(%s)" % b))
    for lname, arrs in own_p.items():
        for bi, w in enumerate(arrs):
            g = net.params[lname][bi].diff.copy()
            w = w.reshape(g.shape)
            scale = max(np.abs(w).max(), 1e-30)
            if bi == 1 and not np.any(g) and np.abs(w).max() < 2e-2 * max(np.abs(own_p[lname][0]).max(), 1e-30):
                # bias of a convolution whose only reader is a batch-statistics BN: the true gradient is 0 (BN removes the
                # mean), the reference accumulates the rounding noise of sum(dx), the device skips that column-sum pass
                e = 0.0
            elif np.abs(w).max() < 1e-4 * max(np.abs(own_p[lname][0]).max(), 1e-30) or np.abs(w).max() == 0:
                e = float(np.abs(g - w).max() / max(np.abs(own_p[lname][0]).max(), 1e-12))   # (zero-gradient blobs: biases in front of a BN, running statistics)
            else:
                e = rel_l2(g, w)
            fr = rel_l2(g, free_p[lname][bi].reshape(g.shape)) if np.abs(free_p[lname][bi]).max() > 0 else 0.0
            report.append(("%s[%d]" % (lname, bi), e, fr))
            if not e <= TOL_STEP_PARAM:
                failures.append(describe_mismatch(g, w, "d %s[%d] (scale %.3g)" % (lname, bi, scale)))
    # forward blobs, teacher-forced (bf16 storage: 1 ulp) -- the TRAIN-phase forward is checked like the TEST-phase one
    for b, g in dev.items():
        if b in inputs or b in skip_data or b == "loss" or b not in forced:
            continue
        e = rel_l2(g, forced[b].reshape(g.shape))
        report.append(("data(%s)" % b, e, rel_l2(g, want[b].reshape(g.shape))))
        if not e <= 4e-3:
            failures.append(describe_mismatch(g, forced[b].reshape(g.shape), "data(%s)" % b))
    # running statistics moved like the reference's: (1 - m) * batch + m * running, from the device's own BN input
    for lname, ps in bn_after.items():
        for k in (2, 3):
            e = rel_max(net.params[lname][k].data.ravel(), ps[k].ravel())
            if not e <= 2e-3:
                failures.append("%s running blob %d rel_max %.3e" % (lname, k, e))
    print("%-40s %-12s %s" % ("blob / parameter", "teacher-forced", "free-running (fp32 oracle end to end)"))
    print("\n".join("%-40s %.3e    %.3e" % r for r in report))
    assert not failures, "\n".join(failures)
    return net, report


def test_toy_train_net_every_gradient(gpu):
    rng = np.random.default_rng(11)
    x = rng.normal(size=(4, 16, 33, 33)).astype(np.float32)
    lab = np.array([1, 7], np.float32).reshape(2, 1, 1, 1)  # the Reshape folds pairs of frames into one clip
    run_train_net_against_oracle(TOY, {"data": x, "label": lab}, 21, "drop", "gp_r", skip_data=("gp", "gp_r"))


def test_eco_lite_train_n4(gpu):
    """BASELINE config #4's network (ECO-Lite train net, Kinetics head) at N=4, two clips: loss, every blob's data and
    gradient, every parameter gradient, BN running statistics."""
    segments, batch, classes = 4, 2, 400
    txt = gen.eco_lite_train(segments=segments, classes=classes, batch=batch)
    x = refnet.eco_input(batch, segments).reshape(batch, 3 * segments, 224, 224)
    lab = np.array([17, 311], np.float32).reshape(batch, 1, 1, 1)
    run_train_net_against_oracle(txt, {"data": x, "label": lab}, 4321, "dropout", "global_pool_reshape",
                                 skip_data=("global_pool", "global_pool_reshape", "reshape_data"))


def test_test_phase_loss_and_accuracy(gpu):
    """the TEST phase of the train/test definition: loss, top-1 and top-5 through the fast plan"""
    import caffe
    segments, batch, classes = 4, 2, 12
    txt = gen.eco_lite_train(segments=segments, classes=classes, batch=batch)
    ref = refnet.RefNet(txt, phase="TEST").init_params(4321)
    x = refnet.eco_input(batch, segments).reshape(batch, 3 * segments, 224, 224)
    ref.calibrate_bn({"data": x, "label": np.zeros((batch, 1, 1, 1), np.float32)})
    net = caffe.Net.from_string(txt, caffe.TEST)
    load_params(net, ref.params_dict())
    want = ref.forward({"data": x, "label": np.zeros((batch, 1, 1, 1), np.float32)}, bf16=True)
    order = np.argsort(-want["fc8"], axis=1)
    lab = np.array([order[0, 0], order[1, 3]], np.float32).reshape(batch, 1, 1, 1)  # clip 0: top-1 hit; clip 1: rank 4
    want = ref.forward({"data": x, "label": lab}, bf16=True)
    net.blobs["data"].data[...] = x
    net.blobs["label"].data[...] = lab
    out = net.forward()
    assert set(out.keys()) == {"loss", "top1", "top5"}
    assert abs(float(out["loss"]) - float(want["loss"])) <= 2e-2 * max(1.0, float(want["loss"]))
    assert float(out["top1"]) == float(want["top1"]) == 0.5
    assert float(out["top5"]) == float(want["top5"]) == 1.0


def test_backward_is_reproducible_run_to_run(gpu):
    """Statistics, dgrad and the elementwise backward kernels are order-deterministic (per-block partials summed in block
    order, no atomics on the activation path): two passes over the same batch give bit-identical blob gradients; only the
    split-K weight gradient adds its partial sums with fp32 atomics (order varies at the 1e-6 level)."""
    segments, batch, classes = 4, 2, 20
    txt = gen.eco_lite_train(segments=segments, classes=classes, batch=batch).replace("dropout_ratio: 0.3", "dropout_ratio: 0")
    net = train_net(txt)
    load_params(net, refnet.RefNet(txt, phase="TRAIN").init_params(4321).params_dict())
    x = refnet.eco_input(batch, segments).reshape(batch, 3 * segments, 224, 224)
    lab = np.array([3, 11], np.float32).reshape(batch, 1, 1, 1)
    runs = []
    for _ in range(2):
        net.blobs["data"].data[...] = x
        net.blobs["label"].data[...] = lab
        loss = float(net.forward()["loss"])
        net.clear_param_diffs()
        net.backward()
        runs.append((loss, net.blobs["conv1_7x7_s2"].diff.copy(), net.blobs["res3a"].diff.copy(),
                     net.params["conv1_7x7_s2"][0].diff.copy(), net.params["res5b_2"][0].diff.copy(),
                     net.params["res3a_bn"][0].diff.copy()))
    a, b = runs
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), "blob gradients differ between identical passes"
    assert np.array_equal(a[5], b[5])
    assert rel_max(a[3], b[3]) <= 1e-5 and rel_max(a[4], b[4]) <= 1e-5


def test_eco_full_train_n4(gpu):
    """the two-stream ECO-Full train net (models_ECO_Full/kinetics/ECO_full.prototxt geometry) at N=4, two clips: 2-D stream
    inception_3c..5b with stride-2 convolutions and MAX pools, segment-consensus pooling and Concat on plain blobs, two
    Dropout layers -- every blob's data and gradient, every parameter gradient (teacher-forced)"""
    segments, batch, classes = 4, 2, 30
    txt = gen.eco_full_train(segments=segments, classes=classes, batch=batch, dropout2d=0.0)
    x = refnet.eco_input(batch, segments).reshape(batch, 3 * segments, 224, 224)
    lab = np.array([7, 21], np.float32).reshape(batch, 1, 1, 1)
    run_train_net_against_oracle(txt, {"data": x, "label": lab}, 4321, "dropout", "global_pool_reshape",
                                 skip_data=("global_pool", "global_pool_reshape", "reshape_data", "global_pool_gn02_reshape"),
                                 skip_diff=("res2b_bn_pre", "res2b_bn"))
