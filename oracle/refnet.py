"""oracle/refnet.py -- TEST INFRASTRUCTURE (the parity checker), not product code.

Runs a caffe_3d net definition on the CPU, layer by layer, through the C
restatement in oracle/ref_cpu.c.  It follows the reference's graph rules:

  * Net::Init / FilterNet phase rules   caffe_3d/src/caffe/net.cpp:39-316, :319-346
  * input / input_dim net inputs          net.cpp:55-73
  * in-place layers share the blob        (top name == bottom name)
  * InsertSplits naming                   caffe_3d/src/caffe/util/insert_splits.cpp:13-142
  * Reshape rules (0 = copy, -1 = infer)  caffe_3d/src/caffe/layers/reshape_layer.cpp:9-90
  * Concat axis 1, Eltwise SUM defaults   caffe.proto:488, :619
  * Dropout TEST = identity               dropout_layer.cpp:46-48

Two numeric modes:
  fp32        -- the reference's arithmetic (fp32 everywhere).
  bf16 mirror -- rounds to bfloat16 at exactly the points where the device path
                 stores bf16 (see DESIGN.md "Rounding contract"), accumulating in
                 fp32 like the tensor cores do, so that the CUDA path can be held
                 to 1e-3 against it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from . import prototxt as _pt

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_lib(force=False):
    so = os.path.join(_HERE, "libref_cpu.so")
    src = os.path.join(_HERE, "ref_cpu.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libref_cpu.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_lib())
        _LIB.ref_num_threads.restype = C.c_int
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(seq):
    return (C.c_int * len(seq))(*[int(v) for v in seq])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------
# thin numpy wrappers over the C functions
def round_bf16(a):
    a = _f32(a)
    out = np.empty_like(a)
    lib().ref_round_bf16(_fp(a), _fp(out), C.c_long(a.size))
    return out


def conv_out_shape(in_sp, kernel, stride, pad):
    return [(i + 2 * p - k) // s + 1 for i, k, s, p in zip(in_sp, kernel, stride, pad)]


def pool_out_shape(in_sp, kernel, stride, pad):
    return [int(lib().ref_pool_out_dim(int(i), int(k), int(s), int(p)))
            for i, k, s, p in zip(in_sp, kernel, stride, pad)]


def conv_forward(x, w, b, kernel, stride, pad, naive=False):
    x, w = _f32(x), _f32(w)
    num, cin = x.shape[:2]
    in_sp = list(x.shape[2:])
    cout = w.shape[0]
    assert w.shape[1] == cin and list(w.shape[2:]) == list(kernel), (w.shape, x.shape, kernel)
    out_sp = conv_out_shape(in_sp, kernel, stride, pad)
    y = np.empty([num, cout] + out_sp, np.float32)
    bp = _fp(_f32(b)) if b is not None else None
    fn = lib().ref_conv_forward_naive if naive else lib().ref_conv_forward
    rc = fn(_fp(x), _fp(w), bp, _fp(y), num, cin, cout, len(in_sp), _ip(in_sp), _ip(kernel),
            _ip(stride), _ip(pad))
    if rc != 0:
        raise RuntimeError("ref_conv_forward rc=%d" % rc)
    return y


def bn_forward_test(x, slope, bias, mean, var, eps=1e-5):
    x = _f32(x)
    y = np.empty_like(x)
    num, ch = x.shape[:2]
    spatial = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    lib().ref_bn_forward_test(_fp(x), _fp(y), _fp(_f32(slope).ravel()), _fp(_f32(bias).ravel()),
                              _fp(_f32(mean).ravel()), _fp(_f32(var).ravel()), C.c_float(eps),
                              num, ch, C.c_long(spatial))
    return y


def bn_forward_train(x, slope, bias, run_mean, run_var, momentum=0.9, eps=1e-5):
    """Returns (y, batch_mean, batch_var); run_mean/run_var (float32 arrays) are updated in place."""
    x = _f32(x)
    y = np.empty_like(x)
    num, ch = x.shape[:2]
    spatial = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    bm = np.empty(ch, np.float32)
    bv = np.empty(ch, np.float32)
    assert run_mean.dtype == np.float32 and run_var.dtype == np.float32
    lib().ref_bn_forward_train(_fp(x), _fp(y), _fp(_f32(slope).ravel()), _fp(_f32(bias).ravel()),
                               _fp(run_mean), _fp(run_var), C.c_float(momentum), C.c_float(eps),
                               num, ch, C.c_long(spatial), _fp(bm), _fp(bv))
    return y, bm, bv


def conv_backward(x, w, dy, kernel, stride, pad, need_dx=True, has_bias=True):
    """Returns (dx or None, dw, db or None) -- conv_layer.cpp:44-75 over base_conv_layer.cpp:290-328."""
    x, w, dy = _f32(x), _f32(w), _f32(dy)
    num, cin = x.shape[:2]
    cout = w.shape[0]
    dx = np.zeros_like(x) if need_dx else None
    dw = np.zeros_like(w)
    db = np.zeros(cout, np.float32) if has_bias else None
    rc = lib().ref_conv_backward(_fp(x), _fp(w), _fp(dy), _fp(dx) if need_dx else None, _fp(dw),
                                 _fp(db) if has_bias else None, num, cin, cout, x.ndim - 2, _ip(x.shape[2:]),
                                 _ip(kernel), _ip(stride), _ip(pad))
    if rc != 0:
        raise RuntimeError("ref_conv_backward rc=%d" % rc)
    return dx, dw, db


def pool_backward(x, dy, kernel, stride, pad, method):
    x, dy = _f32(x), _f32(dy)
    dx = np.empty_like(x)
    rc = lib().ref_pool_backward(_fp(x), _fp(dy), _fp(dx), x.shape[0], x.shape[1], x.ndim - 2, _ip(x.shape[2:]),
                                 _ip(kernel), _ip(stride), _ip(pad), 0 if method == "MAX" else 1)
    if rc != 0:
        raise RuntimeError("ref_pool_backward rc=%d" % rc)
    return dx


def bn_backward_train(x, dy, slope, batch_mean, batch_var, eps=1e-5, need_dx=True):
    """Returns (dx or None, dslope, dbias) -- bn_layer.cpp:241-335."""
    x, dy = _f32(x), _f32(dy)
    num, ch = x.shape[:2]
    spatial = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    dx = np.empty_like(x) if need_dx else None
    ds = np.zeros(ch, np.float32)
    dbias = np.zeros(ch, np.float32)
    lib().ref_bn_backward_train(_fp(x), _fp(dy), _fp(dx) if need_dx else None, _fp(_f32(slope).ravel()),
                                _fp(_f32(batch_mean).ravel()), _fp(_f32(batch_var).ravel()), C.c_float(eps),
                                num, ch, C.c_long(spatial), _fp(ds), _fp(dbias))
    return dx, ds, dbias


def relu(x, slope=0.0):
    x = _f32(x)
    y = np.empty_like(x)
    lib().ref_relu(_fp(x), _fp(y), C.c_long(x.size), C.c_float(slope))
    return y


def pool_forward(x, kernel, stride, pad, method):
    x = _f32(x)
    num, ch = x.shape[:2]
    in_sp = list(x.shape[2:])
    out_sp = pool_out_shape(in_sp, kernel, stride, pad)
    y = np.empty([num, ch] + out_sp, np.float32)
    rc = lib().ref_pool_forward(_fp(x), _fp(y), num, ch, len(in_sp), _ip(in_sp), _ip(kernel),
                                _ip(stride), _ip(pad), 0 if method == "MAX" else 1)
    if rc != 0:
        raise RuntimeError("ref_pool_forward rc=%d" % rc)
    return y


def permute(x, order):
    x = _f32(x)
    y = np.empty([x.shape[o] for o in order], np.float32)
    lib().ref_permute(_fp(x), _fp(y), x.ndim, _ip(x.shape), _ip(order))
    return y


def inner_product(x, w, b):
    x, w = _f32(x), _f32(w)
    m = x.shape[0]
    x2 = x.reshape(m, -1)
    n, k = w.shape[0], int(np.prod(w.shape[1:]))
    assert x2.shape[1] == k, (x.shape, w.shape)
    y = np.empty((m, n), np.float32)
    lib().ref_inner_product(_fp(x2), _fp(w), _fp(_f32(b)) if b is not None else None, _fp(y), m, n, k)
    return y


def eltwise_sum(a, b, ca=1.0, cb=1.0):
    a, b = _f32(a), _f32(b)
    y = np.empty_like(a)
    lib().ref_eltwise_sum(_fp(a), _fp(b), C.c_float(ca), C.c_float(cb), _fp(y), C.c_long(a.size))
    return y


# ----------------------------------------------------------------------------
def _nd_param(p, key, nsp, default, hw=None):
    """kernel_size / stride / pad: given once or once per spatial axis
    (base_conv_layer.cpp:13-110, pooling_layer.cpp:17-114); 2-D nets may use *_h/*_w."""
    if hw is not None and (p.has(hw + "_h") or p.has(hw + "_w")):
        assert nsp == 2
        return [int(p.get1(hw + "_h")), int(p.get1(hw + "_w"))]
    v = p.getall(key)
    if len(v) == 0:
        return [default] * nsp
    if len(v) == 1:
        return [int(v[0])] * nsp
    assert len(v) == nsp, (key, v, nsp)
    return [int(t) for t in v]


def _phase_ok(layer, phase):
    """FilterNet / StateMeetsRule restricted to `phase` rules (net.cpp:319-346, :349-411)."""
    inc, exc = layer.getall("include"), layer.getall("exclude")
    if inc:
        return any((not r.has("phase")) or r.get1("phase") == phase for r in inc)
    return not any(r.has("phase") and r.get1("phase") == phase for r in exc)


class RefLayer:
    def __init__(self, msg):
        self.msg = msg
        self.name = msg.get1("name")
        self.type = msg.get1("type")
        self.bottoms = list(msg.getall("bottom"))
        self.tops = list(msg.getall("top"))
        self.params = []  # list of np arrays in caffe blob order/shape


DATA_TYPES = ("VideoData", "Data", "ImageData", "MemoryData", "Input", "DummyData")


class RefNet:
    def __init__(self, text, phase="TEST", input_shapes=None):
        self.proto = _pt.parse(text)
        self.phase = phase
        self.name = self.proto.get1("name", "")
        self.layers = [RefLayer(m) for m in self.proto.getall("layer") if _phase_ok(m, phase)]
        self.inputs = list(self.proto.getall("input"))
        dims = [int(d) for d in self.proto.getall("input_dim")]
        self.input_shapes = {}
        for i, n in enumerate(self.inputs):
            if self.proto.has("input_shape"):
                self.input_shapes[n] = [int(d) for d in self.proto.getall("input_shape")[i].getall("dim")]
            else:
                self.input_shapes[n] = dims[4 * i:4 * i + 4]
        if input_shapes:
            self.input_shapes.update({k: list(v) for k, v in input_shapes.items()})
        self.blobs = {}
        self._by_name = {l.name: l for l in self.layers}

    # ---- naming: what Net::Init would report after InsertSplits ---------------
    def split_names(self):
        """(layer_names, blob_names) as caffe's Net exposes them, including the
        automatically inserted Split layers (insert_splits.cpp:13-142).  A blob
        produced once and consumed k>1 times gets layer `<blob>_<producer>_<topidx>_split`
        with tops `<blob>_<producer>_<topidx>_split_<j>`."""
        produced = {}   # blob -> (producer layer name, top idx)
        use_count = {}  # (producer, topidx) -> count
        order = []      # producer keys in creation order
        for n in self.inputs:
            produced[n] = ("input", self.inputs.index(n))
        for l in self.layers:
            for b in l.bottoms:
                if b not in produced:
                    raise ValueError("unknown bottom %s" % b)
                key = produced[b] + (b,)
                use_count[key] = use_count.get(key, 0) + 1
            for j, t in enumerate(l.tops):
                produced[t] = (l.name, j)
        # net outputs with loss weights count as a use as well; not needed for ECO deploy
        layer_names, blob_names = [], []

        def add_blob(b):
            if b not in blob_names:
                blob_names.append(b)

        def split_name(b, prod, j):
            return "%s_%s_%d_split" % (b, prod, j)

        for i, n in enumerate(self.inputs):
            add_blob(n)
        for i, n in enumerate(self.inputs):
            key = ("input", i, n)
            if use_count.get(key, 0) > 1:
                sn = split_name(n, "input", i)
                layer_names.append(sn)
                for j in range(use_count[key]):
                    add_blob("%s_%d" % (sn, j))
        produced = {n: ("input", i) for i, n in enumerate(self.inputs)}
        seen_use = {}
        for l in self.layers:
            layer_names.append(l.name)
            for b in l.bottoms:
                pass
            for j, t in enumerate(l.tops):
                produced[t] = (l.name, j)
                add_blob(t)
            for j, t in enumerate(l.tops):
                key = (l.name, j, t)
                if use_count.get(key, 0) > 1:
                    sn = split_name(t, l.name, j)
                    layer_names.append(sn)
                    for q in range(use_count[key]):
                        add_blob("%s_%d" % (sn, q))
        return layer_names, blob_names

    # ---- parameters ------------------------------------------------------------
    def param_shapes(self, batch_shapes=None):
        """layer name -> list of blob shapes, from a shape-only pass."""
        shapes = {}
        self._walk(None, shapes_only=True, out_param_shapes=shapes)
        return shapes

    def init_params(self, seed=4321):
        """Harness-owned weights: caffe's fillers cannot be
        reproduced bit-for-bit and random-init BN running stats of 0/0 overflow, so the
        harness writes every blob: conv/fc W ~ U(+-sqrt(3/fan_in)) (xavier,
        include/caffe/filler.hpp:149-163), bias ~ N(0, 0.01), BN slope ~ U(0.5,1.5),
        bias ~ N(0,0.1), running mean ~ N(0,0.1), running variance ~ U(0.5,1.5)."""
        rng = np.random.default_rng(seed)
        for name, shp in self.param_shapes().items():
            l = self._by_name[name]
            if l.type in ("Convolution", "InnerProduct"):
                fan_in = int(np.prod(shp[0][1:]))
                a = np.sqrt(3.0 / fan_in)
                l.params = [rng.uniform(-a, a, shp[0]).astype(np.float32)]
                if len(shp) > 1:
                    l.params.append(rng.normal(0, 0.01, shp[1]).astype(np.float32))
            elif l.type == "BN":
                l.params = [rng.uniform(0.5, 1.5, shp[0]).astype(np.float32),
                            rng.normal(0, 0.1, shp[1]).astype(np.float32),
                            rng.normal(0, 0.1, shp[2]).astype(np.float32),
                            rng.uniform(0.5, 1.5, shp[3]).astype(np.float32)]
        return self

    def calibrate_bn(self, data, seed=99, jitter=True):
        """Make the harness weights behave like a trained net: walk the net once and set
        every BN layer's running mean/variance to the statistics of its actual input
        (times a small random jitter), so activations stay O(1) through all 30 BN
        layers instead of drifting to 0 or inf.  Purely a weight-generation step."""
        rng = np.random.default_rng(seed)

        def hook(layer, x):
            ch = x.shape[1]
            xs = np.moveaxis(x, 1, 0).reshape(ch, -1).astype(np.float64)
            m, v = xs.mean(1), xs.var(1)
            if jitter:
                m = m + rng.normal(0, 0.1, ch) * np.sqrt(v + 1e-5)
                v = v * rng.uniform(0.8, 1.25, ch)
            layer.params[2] = m.astype(np.float32).reshape(layer.params[2].shape)
            layer.params[3] = np.maximum(v, 1e-4).astype(np.float32).reshape(layer.params[3].shape)

        self._walk({self.inputs[0]: data} if not isinstance(data, dict) else data, bn_hook=hook)
        return self

    def params_dict(self):
        return {l.name: [p.copy() for p in l.params] for l in self.layers if l.params}

    def set_params(self, d):
        for name, arrs in d.items():
            if name in self._by_name:
                self._by_name[name].params = [np.asarray(a, np.float32) for a in arrs]
        return self

    # ---- forward ---------------------------------------------------------------
    def forward(self, inputs, bf16=False, keep=None, teacher=None, teacher_raw=None, dropout_masks=None, seed=7):
        """inputs: array for the single net input or {name: array}.  Returns {blob: array}
        for every blob (in-place layers overwrite, as in caffe).

        teacher: optional {blob: array} (e.g. the device's blobs).  After the last layer that
        writes a blob, the oracle's own value is recorded in the result and the teacher's value is
        substituted for all later consumers, so every layer is checked on the inputs the device
        actually saw (no error amplification through the depth of the net).
        teacher_raw: optional {blob: array} of raw conv / eltwise sums as the device stored them
        (bf16); used only as the *older* operand of an Eltwise, which is exactly what the device's
        fused residual add reads back from memory."""
        if not isinstance(inputs, dict):
            inputs = {self.inputs[0]: inputs}
        self._dropout_masks = dropout_masks or {}
        self._rng = np.random.default_rng(seed)
        return self._walk(inputs, bf16=bf16, teacher=teacher, teacher_raw=teacher_raw)

    def _walk(self, inputs, shapes_only=False, out_param_shapes=None, bf16=False, bn_hook=None, teacher=None,
              teacher_raw=None):
        blobs = {}
        shp = {}
        own = {}
        self.tape = []       # (layer, [bottom arrays], extra) per executed layer, for backward()
        self.bn_batch = {}   # TRAIN-phase BN: layer name -> (batch mean, batch variance)
        plain = set()  # blobs the device keeps as plain fp32 vectors (downstream of a collapsing pool / fc)
        last_writer = {}
        for li, l in enumerate(self.layers):
            for top in l.tops:
                last_writer[top] = li
        R = round_bf16 if bf16 else (lambda a: a)
        if shapes_only:
            for n in self.inputs:
                shp[n] = list(self.input_shapes[n])
        else:
            for n, a in inputs.items():
                blobs[n] = R(_f32(a))
                shp[n] = list(blobs[n].shape)
        produced_order = {n: -1 for n in shp}
        for li, l in enumerate(self.layers):
            t = l.type
            if t in DATA_TYPES:
                for top in l.tops:  # data layers are fed from `inputs`
                    if not shapes_only and top not in blobs:
                        raise ValueError("data layer top %r must be supplied as an input" % top)
                    if shapes_only and top not in shp:
                        shp[top] = list(self.input_shapes.get(top, [1]))
                continue
            bs = [shp[b] for b in l.bottoms]
            extra = {}
            if not shapes_only:
                self.tape.append((l, [blobs[b] for b in l.bottoms], extra))
            if t == "Convolution":
                p = l.msg.get1("convolution_param")
                nsp = len(bs[0]) - 2
                k = _nd_param(p, "kernel_size", nsp, None, "kernel")
                s = _nd_param(p, "stride", nsp, 1, "stride")
                pd = _nd_param(p, "pad", nsp, 0, "pad")
                nout = int(p.get1("num_output"))
                bias_term = bool(p.get1("bias_term", True))
                assert int(p.get1("group", 1)) == 1 and int(p.get1("dilation", 1)) == 1
                osh = bs[0][:1] + [nout] + conv_out_shape(bs[0][2:], k, s, pd)
                if out_param_shapes is not None:
                    out_param_shapes[l.name] = [[nout, bs[0][1]] + k] + ([[nout]] if bias_term else [])
                if not shapes_only:
                    w = R(l.params[0])
                    b = l.params[1] if bias_term else None
                    extra.update(k=k, s=s, pd=pd, bias=bias_term)
                    blobs[l.tops[0]] = conv_forward(blobs[l.bottoms[0]], w, b, k, s, pd)
                shp[l.tops[0]] = osh
            elif t == "BN":
                if out_param_shapes is not None:
                    out_param_shapes[l.name] = [[1, bs[0][1]]] * 4
                if not shapes_only:
                    x = blobs[l.bottoms[0]]
                    if bn_hook is not None:
                        bn_hook(l, x)
                    bp = l.msg.get1("bn_param") or _pt.Msg()
                    eps = float(bp.get1("eps", 1e-5))
                    frozen = bool(bp.get1("frozen", False))
                    if self.phase == "TRAIN" and not frozen:
                        # batch statistics + running-average update (bn_layer.cpp:107-157); the running blobs
                        # of this RefNet are updated in place like the reference's blobs_[2], blobs_[3]
                        rm = np.ascontiguousarray(l.params[2], np.float32).reshape(-1).copy()
                        rv = np.ascontiguousarray(l.params[3], np.float32).reshape(-1).copy()
                        y, bm, bv = bn_forward_train(x, l.params[0], l.params[1], rm, rv,
                                                     float(bp.get1("momentum", 0.9)), eps)
                        l.params[2] = rm.reshape(l.params[2].shape)
                        l.params[3] = rv.reshape(l.params[3].shape)
                        self.bn_batch[l.name] = (bm, bv)
                        extra["train"] = True
                    else:
                        y = bn_forward_test(x, l.params[0], l.params[1], l.params[2], l.params[3], eps)
                    # bf16 mirror: BN output is stored bf16 unless an in-place ReLU follows
                    # (then the rounding happens after the ReLU -- one fused epilogue).
                    nxt = self.layers[li + 1] if li + 1 < len(self.layers) else None
                    fused_relu = nxt is not None and nxt.type == "ReLU" and nxt.bottoms == l.tops and nxt.tops == l.tops
                    blobs[l.tops[0]] = y if fused_relu else R(y)
                shp[l.tops[0]] = list(bs[0])
            elif t == "ReLU":
                if not shapes_only:
                    slope = float((l.msg.get1("relu_param") or _pt.Msg()).get1("negative_slope", 0.0))
                    blobs[l.tops[0]] = R(relu(blobs[l.bottoms[0]], slope))
                shp[l.tops[0]] = list(bs[0])
            elif t == "Pooling":
                p = l.msg.get1("pooling_param")
                nsp = len(bs[0]) - 2
                if bool(p.get1("global_pooling", False)):
                    k = bs[0][2:]
                else:
                    k = _nd_param(p, "kernel_size", nsp, None, "kernel")
                s = _nd_param(p, "stride", nsp, 1, "stride")
                pd = _nd_param(p, "pad", nsp, 0, "pad")
                method = p.get1("pool", "MAX")
                osh = bs[0][:2] + pool_out_shape(bs[0][2:], k, s, pd)
                if not shapes_only:
                    extra.update(k=k, s=s, pd=pd, method=method)
                    y = pool_forward(blobs[l.bottoms[0]], k, s, pd, method)
                    # bf16 mirror: pooled maps stay bf16 feature maps; pools that collapse the
                    # whole map (global pools / segment consensus) feed fp32 vectors to the fc.
                    collapses = all(o == 1 for o in osh[2:])
                    if collapses or l.bottoms[0] in plain:
                        plain.add(l.tops[0])
                    blobs[l.tops[0]] = y if l.tops[0] in plain else R(y)
                shp[l.tops[0]] = osh
            elif t == "Concat":
                axis = int((l.msg.get1("concat_param") or _pt.Msg()).get1("axis", 1))
                osh = list(bs[0])
                osh[axis] = sum(b[axis] for b in bs)
                if not shapes_only:
                    extra["axis"] = axis
                    blobs[l.tops[0]] = np.concatenate([blobs[b] for b in l.bottoms], axis=axis)
                shp[l.tops[0]] = osh
            elif t == "Eltwise":
                ep = l.msg.get1("eltwise_param") or _pt.Msg()
                assert ep.get1("operation", "SUM") == "SUM"
                co = [float(c) for c in ep.getall("coeff")] or [1.0] * len(l.bottoms)
                assert len(l.bottoms) == 2
                if not shapes_only:
                    a, b = blobs[l.bottoms[0]], blobs[l.bottoms[1]]
                    if bf16:
                        # the bottom produced last is still in the fp32 accumulator when the
                        # add happens (fused into that conv's epilogue); the other one is read
                        # back from its bf16 copy in HBM.
                        last = max(range(2), key=lambda i: produced_order[l.bottoms[i]])
                        older = l.bottoms[1 - last]
                        if teacher_raw is not None and older in teacher_raw:
                            old_val = _f32(teacher_raw[older]).reshape(a.shape)
                        else:
                            old_val = round_bf16(b if last == 0 else a)
                        if last == 0:
                            b = old_val
                        else:
                            a = old_val
                    extra["coeff"] = co
                    blobs[l.tops[0]] = eltwise_sum(a, b, co[0], co[1])
                shp[l.tops[0]] = list(bs[0])
            elif t == "Reshape":
                dims = [int(d) for d in l.msg.get1("reshape_param").get1("shape").getall("dim")]
                rp = l.msg.get1("reshape_param")
                assert int(rp.get1("axis", 0)) == 0 and int(rp.get1("num_axes", -1)) == -1
                osh = []
                for i, d in enumerate(dims):
                    osh.append(bs[0][i] if d == 0 else d)
                cnt = int(np.prod(bs[0]))
                if -1 in osh:
                    known = int(np.prod([d for d in osh if d != -1]))
                    osh[osh.index(-1)] = cnt // known
                assert int(np.prod(osh)) == cnt, (l.name, bs[0], osh)
                if not shapes_only:
                    blobs[l.tops[0]] = blobs[l.bottoms[0]].reshape(osh)
                shp[l.tops[0]] = osh
            elif t == "Permute":
                order = [int(o) for o in l.msg.get1("permute_param").getall("order")]
                order += [i for i in range(len(bs[0])) if i not in order]
                if not shapes_only:
                    extra["order"] = order
                    blobs[l.tops[0]] = permute(blobs[l.bottoms[0]], order)
                shp[l.tops[0]] = [bs[0][o] for o in order]
            elif t == "Dropout":
                if not shapes_only:
                    if self.phase == "TEST":
                        blobs[l.tops[0]] = blobs[l.bottoms[0]]
                    else:
                        # dropout_layer.cpp:33-49: Bernoulli(1-p) mask, survivors scaled by 1/(1-p).  caffe's RNG
                        # stream is not reproducible outside caffe; a mask can be supplied (e.g. the device's)
                        ratio = float((l.msg.get1("dropout_param") or _pt.Msg()).get1("dropout_ratio", 0.5))
                        xin = blobs[l.bottoms[0]]
                        mask = self._dropout_masks.get(l.name)
                        if mask is None:
                            mask = (self._rng.random(xin.shape) < (1.0 - ratio))
                        mask = np.asarray(mask, np.float32).reshape(xin.shape)
                        extra["mask"], extra["scale"] = mask, np.float32(1.0 / (1.0 - ratio))
                        blobs[l.tops[0]] = (xin * mask * extra["scale"]).astype(np.float32)
                shp[l.tops[0]] = list(bs[0])
            elif t == "InnerProduct":
                p = l.msg.get1("inner_product_param")
                nout = int(p.get1("num_output"))
                bias_term = bool(p.get1("bias_term", True))
                k = int(np.prod(bs[0][1:]))
                if out_param_shapes is not None:
                    out_param_shapes[l.name] = [[nout, k]] + ([[nout]] if bias_term else [])
                if not shapes_only:
                    blobs[l.tops[0]] = inner_product(blobs[l.bottoms[0]], l.params[0],
                                                     l.params[1] if bias_term else None)
                shp[l.tops[0]] = [bs[0][0], nout]
            elif t == "Softmax":
                if not shapes_only:
                    x = blobs[l.bottoms[0]]
                    e = np.exp(x - x.max(1, keepdims=True))
                    blobs[l.tops[0]] = (e / e.sum(1, keepdims=True)).astype(np.float32)
                shp[l.tops[0]] = list(bs[0])
            elif t in ("SoftmaxWithLoss", "Accuracy"):
                if not shapes_only:
                    x = blobs[l.bottoms[0]].astype(np.float64)
                    lab = blobs[l.bottoms[1]].astype(np.int64).ravel()
                    if t == "SoftmaxWithLoss":
                        # softmax_loss_layer.cpp:48-77: loss = -sum log(max(prob[label], FLT_MIN)) / count (normalize)
                        z = x - x.max(1, keepdims=True)
                        prob = np.exp(z) / np.exp(z).sum(1, keepdims=True)
                        extra["prob"], extra["label"] = prob, lab
                        pl = np.maximum(prob[np.arange(len(lab)), lab], np.finfo(np.float32).tiny)
                        blobs[l.tops[0]] = np.float32(-np.log(pl).sum() / len(lab)).reshape(())
                    else:
                        topk = int((l.msg.get1("accuracy_param") or _pt.Msg()).get1("top_k", 1))
                        # accuracy_layer.cpp: label counted if among the top_k scores
                        idx = np.argsort(-x, axis=1, kind="stable")[:, :topk]
                        blobs[l.tops[0]] = np.float32((idx == lab[:, None]).any(1).mean()).reshape(())
                for top in l.tops:
                    shp[top] = []
            else:
                raise NotImplementedError("oracle: layer type %s (%s)" % (t, l.name))
            if t in ("InnerProduct", "Softmax") or (l.bottoms and l.bottoms[0] in plain and t in
                                                     ("Reshape", "Dropout", "Concat", "Permute", "Split")):
                plain.update(l.tops)
            for top in l.tops:
                produced_order[top] = li
                if teacher is not None and not shapes_only and last_writer[top] == li and top in blobs:
                    own[top] = blobs[top]
                    if top in teacher:
                        blobs[top] = _f32(teacher[top]).reshape(blobs[top].shape)
        self.shapes = shp
        self.blobs = blobs
        if teacher is not None and not shapes_only:
            for k, v in blobs.items():
                own.setdefault(k, v)
            return own
        return blobs


    # ---- backward -----------------------------------------------------------------
    def backward(self, top_diffs=None, loss_weight=1.0, teacher_diffs=None):
        """Backward pass over the tape of the last fp32 forward(), layer by layer in reverse (Net::BackwardFromTo,
        net.cpp:637-706).  Blob diffs accumulate over all consumers of a blob (what the auto-inserted Split
        layers do, split_layer.cpp); in-place layers replace the diff of their blob.  Parameter diffs are returned
        fresh (zero-initialised), i.e. one iteration with iter_size 1.
        top_diffs: optional {blob: dL/dblob} seeds (default: loss layers seed themselves with loss_weight).
        teacher_diffs: optional {blob: dL/dblob as the DEVICE computed it}.  When the pass reaches the last writer of
        such a blob, the oracle's own (accumulated) diff is recorded and the teacher's is substituted for everything
        upstream, so each layer's backward is judged on the top gradient the device actually had (the backward
        analogue of forward(teacher=...)); combine with a teacher-forced forward so the tape holds the device's
        activations.  The recorded own values are what `blob_diffs` returns then.
        Returns (blob_diffs, param_diffs): {blob: array}, {layer: [arrays in blob order]}."""
        own = {}
        last_writer = {}
        for l, _, _ in self.tape:
            for t_ in l.tops:
                last_writer[t_] = l
        diffs = {}
        pdiffs = {}
        data_blobs = set(self.inputs)
        for l in self.layers:
            if l.type in DATA_TYPES:
                data_blobs.update(l.tops)
        needs = set()  # blobs whose gradient is needed: everything downstream of a parameterised layer
        for l, _, _ in self.tape:
            if l.params or any(b in needs for b in l.bottoms):
                needs.update(l.tops)
        if top_diffs:
            for k, v in top_diffs.items():
                diffs[k] = _f32(v).copy()

        def add(name, g, inplace=False):
            if name in data_blobs or name not in needs:
                return
            if inplace or name not in diffs:
                diffs[name] = _f32(g)
            else:
                diffs[name] = diffs[name] + g

        for l, bots, extra in reversed(self.tape):
            t = l.type
            if t in DATA_TYPES or t == "Accuracy":
                continue
            if teacher_diffs is not None:
                for t_ in l.tops:
                    if last_writer.get(t_) is l and t_ in diffs and t_ in teacher_diffs and t_ not in own:
                        own[t_] = diffs[t_]
                        diffs[t_] = _f32(teacher_diffs[t_]).reshape(diffs[t_].shape)
            if t == "SoftmaxWithLoss":
                prob, lab = extra["prob"], extra["label"]
                g = prob.copy()
                g[np.arange(len(lab)), lab] -= 1.0
                w = float(diffs.get(l.tops[0], loss_weight)) if l.tops[0] in diffs else loss_weight
                add(l.bottoms[0], (g * (w / len(lab))).astype(np.float32).reshape(bots[0].shape))
                continue
            top = l.tops[0]
            if top not in diffs:
                continue  # nothing flows back through this layer
            dy = diffs[top]
            inplace = l.bottoms and l.bottoms[0] == top
            x = bots[0] if bots else None
            if t == "Convolution":
                need_dx = l.bottoms[0] in needs and l.bottoms[0] not in data_blobs
                dx, dw, db = conv_backward(x, l.params[0], dy, extra["k"], extra["s"], extra["pd"], need_dx, extra["bias"])
                pdiffs[l.name] = [dw] + ([db] if extra["bias"] else [])
                if need_dx:
                    add(l.bottoms[0], dx)
            elif t == "BN":
                bp = l.msg.get1("bn_param") or _pt.Msg()
                eps = float(bp.get1("eps", 1e-5))
                if extra.get("train"):
                    bm, bv = self.bn_batch[l.name]
                    dx, ds, dbias = bn_backward_train(x, dy, l.params[0], bm, bv, eps)
                    pdiffs[l.name] = [ds.reshape(l.params[0].shape), dbias.reshape(l.params[1].shape),
                                      np.zeros_like(l.params[2]), np.zeros_like(l.params[3])]
                else:  # frozen / TEST statistics: dx = dy * slope / sqrt(var + eps)   (bn_layer.cpp:213-239)
                    sc = (l.params[0].ravel() * np.power(l.params[3].ravel() + np.float32(eps), np.float32(-0.5)))
                    dx = dy * sc.reshape([1, -1] + [1] * (dy.ndim - 2)).astype(np.float32)
                add(l.bottoms[0], dx, inplace)
            elif t == "ReLU":
                slope = float((l.msg.get1("relu_param") or _pt.Msg()).get1("negative_slope", 0.0))
                add(l.bottoms[0], (dy * ((x > 0) + slope * (x <= 0))).astype(np.float32), inplace)  # relu_layer.cpp:24-38
            elif t == "Pooling":
                add(l.bottoms[0], pool_backward(x, dy, extra["k"], extra["s"], extra["pd"], extra["method"]))
            elif t == "Concat":
                off = 0
                for b, arr in zip(l.bottoms, bots):
                    n = arr.shape[extra["axis"]]
                    add(b, np.ascontiguousarray(np.take(dy, range(off, off + n), axis=extra["axis"])))
                    off += n
            elif t == "Eltwise":
                for b, c in zip(l.bottoms, extra["coeff"]):
                    add(b, dy * np.float32(c))
            elif t == "Reshape":
                add(l.bottoms[0], dy.reshape(x.shape))
            elif t == "Permute":
                add(l.bottoms[0], np.ascontiguousarray(np.transpose(dy, np.argsort(extra["order"]))))
            elif t == "Dropout":
                add(l.bottoms[0], dy * extra["mask"] * extra["scale"] if "mask" in extra else dy, inplace)
            elif t == "InnerProduct":
                m = x.shape[0]
                x2, dy2 = x.reshape(m, -1), dy.reshape(m, -1)
                pdiffs[l.name] = [(dy2.T @ x2).astype(np.float32)] + ([dy2.sum(0).astype(np.float32)] if len(l.params) > 1 else [])
                add(l.bottoms[0], (dy2 @ l.params[0]).astype(np.float32).reshape(x.shape))   # inner_product_layer.cpp:96-120
            elif t == "Softmax":
                raise NotImplementedError("oracle backward: plain Softmax is not on the training path")
            else:
                raise NotImplementedError("oracle backward: layer type %s (%s)" % (t, l.name))
        if teacher_diffs is not None:
            for k, v in diffs.items():
                own.setdefault(k, v)
            return own, pdiffs
        return diffs, pdiffs


def eco_input(batch_videos, segments, seed=1234, size=224):
    """Synthetic frames per SURVEY.md 8(d): uint8 U{0..255}, then what the data layer would
    hand over -- fp32, BGR mean (104,117,123) subtracted, no scale.  Shape [B*N,3,H,W]."""
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, size=(batch_videos * segments, 3, size, size), dtype=np.uint8)
    mean = np.array([104.0, 117.0, 123.0], np.float32).reshape(1, 3, 1, 1)
    return u8.astype(np.float32) - mean
