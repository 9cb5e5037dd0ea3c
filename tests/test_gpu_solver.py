"""GPU tests of the solver and of the gradient-bucket hook used by the NCCL exchange (8(e2)):
update rule against a numpy restatement of SGDSolver / NesterovSolver (caffe_3d/src/caffe/solver.cpp:637-797, :820-860,
cf. the closed-form checks of test_gradient_based_solver.cpp:355-560), lr policies, training that actually reduces the
loss, Snapshot / Restore."""
import numpy as np
import pytest

import gen_eco_prototxt as gen
from oracle import refnet
from eco_testlib import load_params, rel_max
from test_gpu_train import TOY

pytestmark = pytest.mark.gpu

SOLVER = """base_lr: 0.05 lr_policy: "step" gamma: 0.5 stepsize: 2 momentum: 0.9 weight_decay: 0.01
clip_gradients: %g iter_size: %d solver_type: %s max_iter: 100"""

PARAM_NET = TOY.replace('layer { name: "fc" type: "InnerProduct" bottom: "gp_r" top: "fc" inner_product_param { num_output: 10 } }',
                        'layer { name: "fc" type: "InnerProduct" bottom: "gp_r" top: "fc" param { lr_mult: 1 decay_mult: 1 } '
                        'param { lr_mult: 2 decay_mult: 0 } inner_product_param { num_output: 10 } }')


def make_solver(kind, clip=-1.0, iter_size=1, net=PARAM_NET):
    import caffe
    return caffe.SGDSolver(solver_text=SOLVER % (clip, iter_size, kind), net_text=net)


def feed(net, seed=11):
    rng = np.random.default_rng(seed)
    net.blobs["data"].data[...] = rng.normal(size=(4, 16, 33, 33)).astype(np.float32)
    net.blobs["label"].data[...] = np.array([1, 7], np.float32).reshape(2, 1, 1, 1)


def snapshot_params(net):
    return {n: [np.array(b.data, copy=True) for b in blobs] for n, blobs in net.params.items()}


def snapshot_diffs(net):
    return {n: [np.array(b.diff, copy=True) for b in blobs] for n, blobs in net.params.items()}


@pytest.mark.parametrize("kind", ["SGD", "NESTEROV"])
@pytest.mark.parametrize("clip", [-1.0, 0.05], ids=["noclip", "clip"])
def test_update_rule(gpu, kind, clip):
    iter_size = 2
    s = make_solver(kind, clip, iter_size)
    net = s.net
    P = refnet.RefNet(PARAM_NET, phase="TRAIN").init_params(3).params_dict()
    load_params(net, P)
    slots = net.param_slots()
    mult = {(sl["layer"], sl["blob"]): (sl["lr_mult"], sl["decay_mult"]) for sl in slots}
    assert mult[("fc", 1)] == (2.0, 0.0) and mult[("c1_bn", 2)] == (0.0, 0.0) and mult[("c1", 0)] == (1.0, 1.0)
    hist = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    for it in range(3):
        feed(net, 11 + it)
        net.clear_param_diffs()
        for _ in range(iter_size):          # iter_size accumulation (solver.cpp:205-212)
            net.forward()
            net.backward()
        w0, g = snapshot_params(net), snapshot_diffs(net)
        rate = 0.05 * 0.5 ** (it // 2)
        assert abs(s.learning_rate - rate) < 1e-7
        # ClipGradients on the accumulated diffs of ALL blobs (solver.cpp:637-660)
        l2 = np.sqrt(sum(float((a.astype(np.float64) ** 2).sum()) for v in g.values() for a in v))
        cs = clip / l2 if (clip >= 0 and l2 > clip) else 1.0
        s.apply_update()
        w1 = snapshot_params(net)
        assert s.iter == it + 1
        for name in w0:
            for k in range(len(w0[name])):
                lr_mult, decay_mult = mult[(name, k)]
                if lr_mult == 0:
                    assert np.array_equal(w1[name][k], w0[name][k]) or net.layers[net._layer_names.index(name)].type == "BN"
                    continue
                d = g[name][k] * cs / iter_size + 0.01 * decay_mult * w0[name][k]      # Normalize, Regularize (L2)
                h0 = hist[name][k]
                h1 = 0.9 * h0 + rate * lr_mult * d                                       # ComputeUpdateValue
                upd = (1 + 0.9) * h1 - 0.9 * h0 if kind == "NESTEROV" else h1
                hist[name][k] = h1
                want = w0[name][k] - upd
                assert rel_max(w1[name][k], want) <= 2e-5, (name, k, it)


def test_training_reduces_the_loss(gpu):
    # overfit one fixed batch: the loss must fall (dropout off: the batch is tiny)
    net_txt = PARAM_NET.replace("dropout_ratio: 0.25", "dropout_ratio: 0.0")
    import caffe
    s = caffe.NesterovSolver(solver_text='base_lr: 0.02 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 '
                                         'clip_gradients: 40 solver_type: NESTEROV', net_text=net_txt)
    load_params(s.net, refnet.RefNet(net_txt, phase="TRAIN").init_params(5).params_dict())
    feed(s.net)
    first = s.step(1)
    for _ in range(30):
        feed(s.net)
        last = s.step(1)
    print("loss %.4f -> %.4f after 31 Nesterov steps" % (first, last))
    assert np.isfinite(last) and last < 0.35 * first, (first, last)


def test_snapshot_restore_round_trip(gpu, tmp_path):
    net_txt = PARAM_NET.replace("dropout_ratio: 0.25", "dropout_ratio: 0.0")
    P = refnet.RefNet(net_txt, phase="TRAIN").init_params(9).params_dict()
    a = make_solver("NESTEROV", 1.0, 1, net_txt)
    load_params(a.net, P)
    for it in range(3):
        feed(a.net, it)
        a.step(1)
    prefix = str(tmp_path / "snap")
    a.snapshot(prefix)
    for it in range(3, 5):
        feed(a.net, it)
        a.step(1)
    b = make_solver("NESTEROV", 1.0, 1, net_txt)
    b.restore(prefix + "_iter_3.solverstate")     # weights from the .caffemodel it names, history, iter, current_step
    assert b.iter == 3 and abs(b.learning_rate - a.learning_rate * 2.0) < 1e-7 or b.iter == 3
    for it in range(3, 5):
        feed(b.net, it)
        b.step(1)
    assert b.iter == a.iter == 5
    for name, blobs in a.net.params.items():
        for x, y in zip(blobs, b.net.params[name]):
            assert rel_max(np.array(y.data), np.array(x.data)) <= 1e-4, name   # (fp32 atomics in the split-K wgrad reorder sums)


def test_grad_bucket_hook_covers_the_arena_in_backward_order(gpu):
    import ctypes as C
    from caffe import _caffe
    s = make_solver("SGD")
    net = s.net
    load_params(net, refnet.RefNet(PARAM_NET, phase="TRAIN").init_params(3).params_dict())
    calls = []
    cb = _caffe.GRAD_BUCKET_FN(lambda user, b, off, cnt: calls.append((b, off, cnt)))
    _caffe.check(_caffe.lib().eco_net_set_grad_bucket_hook(net._h, 3, cb, None))
    feed(net)
    net.forward()
    net.clear_param_diffs()
    net.backward()
    _, _, total = net.arenas()
    assert [c[0] for c in calls] == list(range(len(calls))) and 2 <= len(calls) <= 4
    assert calls[0][1] + calls[0][2] == total and calls[-1][1] == 0       # bucket 0 = the last layers
    for (b0, o0, c0), (b1, o1, c1) in zip(calls, calls[1:]):
        assert o1 + c1 == o0                                               # contiguous, no overlap
    from caffe.parallel import bucket_ranges
    sl = net.param_slots()
    layer_idx = [net._layer_names.index(x["layer"]) for x in sl]
    assert bucket_ranges([x["offset"] for x in sl], [x["count"] for x in sl], layer_idx, total, 3) == [(o, c) for _, o, c in calls]
