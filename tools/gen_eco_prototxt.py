#!/usr/bin/env python
"""Generate ECO-Lite / ECO-Full deploy prototxts for any num_segments / class count.

The reference ships one hand-edited prototxt per dataset with N=16 baked in
(models_ECO_Lite/*/deploy.prototxt, models_ECO_Full/*/deploy.prototxt) and asks
the user to edit three fields to change N (README.md:85-95: `r2Dto3D` dim,
`global_pool` kernel, input_dim).  /root/reference does not exist on the GPU box,
so tests/bench/smoke build their nets from this generator instead; layer names,
blob names, layer order and every parameter that affects the forward pass follow
the reference files (tests/test_prototxt_gen.py checks the parsed structures are
identical when /root/reference is present).

Usage: gen_eco_prototxt.py {lite|full} [--segments N] [--classes C] [--fc NAME]
                           [--batch B] [--dropout R] > deploy.prototxt
"""
from __future__ import annotations

import argparse
import io

DATASETS = {  # dataset -> (fc layer name lite, classes, lite dropout, net name)
    "ucf101": ("fc8u", 101, 0.6, "o3d"),
    "hmdb51": ("fc8h", 51, 0.6, "o3d"),
    "kinetics": ("fc8", 400, 0.3, "ECOLite"),
    "something_something": ("fc8u", 174, 0.3, "o3d"),
}


class _W:
    def __init__(self):
        self.f = io.StringIO()

    def w(self, s=""):
        self.f.write(s + "\n")


def _conv2d(o, name, bottom, nout, k, stride=None, pad=None):
    o.w('layer { name: "%s" type: "Convolution" bottom: "%s" top: "%s"' % (name, bottom, name))
    o.w("  param { lr_mult: 1.0 decay_mult: 1.0 } param { lr_mult: 1.0 decay_mult: 2.0 }")
    o.w("  convolution_param { num_output: %d" % nout)
    if pad is not None:
        o.w("    pad: %d" % pad)
    o.w("    kernel_size: %d" % k)
    if stride is not None:
        o.w("    stride: %d" % stride)
    o.w('    weight_filler { type: "xavier" } bias_filler { type: "constant" value: 0.0 } } }')


def _bn2d(o, name, bottom):
    o.w('layer { name: "%s" type: "BN" bottom: "%s" top: "%s"' % (name, bottom, name))
    o.w("  param { lr_mult: 1.0 decay_mult: 0.0 } param { lr_mult: 1.0 decay_mult: 0.0 }")
    o.w('  bn_param { slope_filler { type: "constant" value: 1.0 } bias_filler { type: "constant" value: 0.0 } } }')


def _relu(o, name, blob):
    o.w('layer { name: "%s" type: "ReLU" bottom: "%s" top: "%s" }' % (name, blob, blob))


def _cbr2d(o, prefix, suffix, bottom, nout, k, stride=None, pad=None):
    """conv `<prefix>_<suffix>` -> BN `<..>_bn` -> in-place ReLU `<prefix>_relu_<suffix>_inp`."""
    name = "%s_%s" % (prefix, suffix)
    _conv2d(o, name, bottom, nout, k, stride, pad)
    _bn2d(o, name + "_bn", name)
    # reference quirk: the stem's ReLU is `conv1_relu_7x7_inp` (no `_s2`)
    _relu(o, "%s_relu_%s_inp" % (prefix, suffix.replace("7x7_s2", "7x7")), name + "_bn")
    return name + "_bn"


def _pool(o, name, bottom, method, k, stride=None, pad=None, top=None):
    o.w('layer { name: "%s" type: "Pooling" bottom: "%s" top: "%s"' % (name, bottom, top or name))
    s = "  pooling_param { pool: %s kernel_size: %d" % (method, k)
    if stride is not None:
        s += " stride: %d" % stride
    if pad is not None:
        s += " pad: %d" % pad
    o.w(s + " } }")


def _concat(o, name, bottoms, top=None, axis=None):
    o.w('layer { name: "%s" type: "Concat" %s top: "%s"%s }' % (
        name, " ".join('bottom: "%s"' % b for b in bottoms), top or name,
        "" if axis is None else " concat_param { axis: %d }" % axis))


def _inception(o, p, bottom, c1, c3r, c3, cd3r, cd3a, cd3b, pool, cproj, stride2=False):
    """One BN-Inception block `inception_<p>`; stride2 blocks have no 1x1/proj branch
    and pass the MAX-pooled input straight into the concat."""
    pre = "inception_" + p
    outs = []
    if not stride2:
        outs.append(_cbr2d(o, pre, "1x1", bottom, c1, 1))
    t = _cbr2d(o, pre, "3x3_reduce", bottom, c3r, 1)
    outs.append(_cbr2d(o, pre, "3x3", t, c3, 3, 2 if stride2 else None, 1))
    t = _cbr2d(o, pre, "double_3x3_reduce", bottom, cd3r, 1)
    t = _cbr2d(o, pre, "double_3x3_1", t, cd3a, 3, None, 1)
    outs.append(_cbr2d(o, pre, "double_3x3_2", t, cd3b, 3, 2 if stride2 else None, 1))
    if stride2:
        _pool(o, pre + "_pool", bottom, "MAX", 3, 2)
        outs.append(pre + "_pool")
    else:
        if pool == "AVE":
            _pool(o, pre + "_pool", bottom, "AVE", 3, 1, 1)
        else:
            _pool(o, pre + "_pool", bottom, "MAX", 3, 1, 1)
        outs.append(_cbr2d(o, pre, "pool_proj", pre + "_pool", cproj, 1))
    _concat(o, pre + "_output", outs)
    return pre + "_output"


def _conv3d(o, name, bottom, top, nout, stride):
    o.w('layer { name: "%s" bottom: "%s" top: "%s" type: "Convolution"' % (name, bottom, top))
    o.w("  convolution_param { num_output: %d pad: [1, 1, 1] kernel_size: [3, 3, 3] stride: [%d, %d, %d]"
        % (nout, stride, stride, stride))
    o.w('    weight_filler { type: "xavier" } bias_filler { type: "constant" value: 0 } }')
    o.w("  param { lr_mult: 1 decay_mult: 1 } param { lr_mult: 2 decay_mult: 0 } }")


def _bn3d(o, name, bottom):
    o.w('layer { name: "%s" bottom: "%s" top: "%s" type: "BN"' % (name, bottom, name))
    o.w("  param { lr_mult: 1 decay_mult: 0 } param { lr_mult: 1 decay_mult: 0 }")
    o.w('  bn_param { frozen: false slope_filler { type: "constant" value: 1 } bias_filler { type: "constant" value: 0 } } }')


def _bnrelu3d(o, stem, bottom):
    _bn3d(o, stem + "_bn", bottom)
    o.w('layer { name: "%s_relu" bottom: "%s_bn" top: "%s_bn" type: "ReLU" }' % (stem, stem, stem))
    return stem + "_bn"


def _eltwise(o, name, a, b):
    o.w('layer { name: "%s" bottom: "%s" bottom: "%s" top: "%s" type: "Eltwise" }' % (name, a, b, name))


def _trunk_to_3c(o, data_blob):
    t = _cbr2d(o, "conv1", "7x7_s2", data_blob, 64, 7, 2, 3)
    _pool(o, "pool1_3x3_s2", t, "MAX", 3, 2)
    t = _cbr2d(o, "conv2", "3x3_reduce", "pool1_3x3_s2", 64, 1)
    t = _cbr2d(o, "conv2", "3x3", t, 192, 3, None, 1)
    _pool(o, "pool2_3x3_s2", t, "MAX", 3, 2)
    t = _inception(o, "3a", "pool2_3x3_s2", 64, 64, 64, 64, 96, 96, "AVE", 32)
    t = _inception(o, "3b", t, 64, 64, 96, 64, 96, 96, "AVE", 64)
    return t


def _head3d(o, bottom2d, segments):
    o.w('layer { name: "r2Dto3D" type: "Reshape" bottom: "%s" top: "res2b_bn_pre"' % bottom2d)
    o.w("  reshape_param { shape { dim: -1 dim: %d dim: 96 dim: 28 dim: 28 } } }" % segments)
    o.w('layer { name: "Transpose1" type: "Permute" bottom: "res2b_bn_pre" top: "res2b_bn"')
    o.w("  permute_param { order: [0,2,1,3,4] } }")
    _conv3d(o, "res3a_2n", "res2b_bn", "res3a", 128, 1)
    t = _bnrelu3d(o, "res3a", "res3a")
    _conv3d(o, "res3b_1", t, "res3b_1", 128, 1)
    t = _bnrelu3d(o, "res3b_1", "res3b_1")
    _conv3d(o, "res3b_2", t, "res3b_2", 128, 1)
    _eltwise(o, "res3b", "res3b_2", "res3a")
    prev_raw, t = "res3b", _bnrelu3d(o, "res3b", "res3b")
    for stage, ch in (("res4", 256), ("res5", 512)):
        a, b = stage + "a", stage + "b"
        _conv3d(o, a + "_1", t, a + "_1", ch, 2)
        u = _bnrelu3d(o, a + "_1", a + "_1")
        _conv3d(o, a + "_2", u, a + "_2", ch, 1)
        _conv3d(o, a + "_down", t, a + "_down", ch, 2)
        _eltwise(o, a, a + "_2", a + "_down")
        t = _bnrelu3d(o, a, a)
        _conv3d(o, b + "_1", t, b + "_1", ch, 1)
        u = _bnrelu3d(o, b + "_1", b + "_1")
        _conv3d(o, b + "_2", u, b + "_2", ch, 1)
        _eltwise(o, b, b + "_2", a)
        t = _bnrelu3d(o, b, b)
    return t  # res5b_bn


def _tail3d(o, segments, dropout):
    assert segments % 4 == 0, "the 3-D head halves the depth twice"
    o.w('layer { name: "global_pool" bottom: "res5b_bn" top: "global_pool" type: "Pooling"')
    o.w("  pooling_param { pool: AVE kernel_size: [%d, 7, 7] stride: [1, 1, 1] } }" % (segments // 4))
    o.w('layer { name: "global_pool_reshape" type: "Reshape" bottom: "global_pool" top: "global_pool_reshape"')
    o.w("  reshape_param { shape { dim: -1 dim: 512 } } }")
    o.w('layer { name: "dropout" type: "Dropout" bottom: "global_pool_reshape" top: "global_pool_reshape"')
    o.w("  dropout_param { dropout_ratio: %g } }" % dropout)


def _fc(o, name, bottom, classes):
    o.w('layer { name: "%s" type: "InnerProduct" bottom: "%s" top: "fc8"' % (name, bottom))
    o.w("  param { lr_mult: 1 decay_mult: 1 } param { lr_mult: 2 decay_mult: 0 }")
    o.w('  inner_product_param { num_output: %d weight_filler { type: "xavier" } bias_filler { type: "constant" value: 0 } } }' % classes)


def _header(o, net_name, frames):
    o.w('name: "%s"' % net_name)
    o.w('input: "data"')
    for d in (frames, 3, 224, 224):
        o.w("input_dim: %d" % d)


def eco_lite_deploy(segments=16, classes=101, fc_name="fc8u", batch=5, dropout=0.6, net_name="o3d"):
    """ECO-Lite deploy net (cf. models_ECO_Lite/ucf101/deploy.prototxt; the reference
    declares input_dim 80 = 5 videos x 16 frames)."""
    o = _W()
    _header(o, net_name, batch * segments)
    t = _trunk_to_3c(o, "data")
    t = _cbr2d(o, "inception_3c", "double_3x3_reduce", t, 64, 1)
    t = _cbr2d(o, "inception_3c", "double_3x3_1", t, 96, 3, None, 1)
    _head3d(o, t, segments)
    _tail3d(o, segments, dropout)
    _fc(o, fc_name, "global_pool_reshape", classes)
    return o.f.getvalue()


def eco_lite_train(segments=16, classes=400, fc_name="fc8", batch=17, dropout=0.3, net_name="o3d"):
    """ECO-Lite train/test net (cf. models_ECO_Lite/kinetics/ECO_Lite.prototxt): the VideoData layer is replaced by the two
    net inputs it would produce -- `data` [B, 3N, 224, 224] and `label` [B, 1, 1, 1] (the GPU data pipeline is SURVEY 8(f1)) --
    then `reshape_data`, the same trunk / head / classifier as the deploy net, SoftmaxWithLoss in both phases and the two
    Accuracy layers in TEST phase."""
    o = _W()
    o.w('name: "%s"' % net_name)
    o.w('input: "data"')
    for d in (batch, 3 * segments, 224, 224):
        o.w("input_dim: %d" % d)
    o.w('input: "label"')
    for d in (batch, 1, 1, 1):
        o.w("input_dim: %d" % d)
    o.w('layer { name: "reshape_data" type: "Reshape" bottom: "data" top: "reshape_data" '
        'reshape_param { shape { dim: -1 dim: 3 dim: 224 dim: 224 } } }')
    t = _trunk_to_3c(o, "reshape_data")
    t = _cbr2d(o, "inception_3c", "double_3x3_reduce", t, 64, 1)
    t = _cbr2d(o, "inception_3c", "double_3x3_1", t, 96, 3, None, 1)
    _head3d(o, t, segments)
    _tail3d(o, segments, dropout)
    _fc(o, fc_name, "global_pool_reshape", classes)
    o.w('layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc8" bottom: "label" include { phase: TRAIN } top: "loss" }')
    o.w('layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc8" bottom: "label" top: "loss" include { phase: TEST } }')
    o.w('layer { name: "top1" type: "Accuracy" bottom: "fc8" bottom: "label" top: "top1" accuracy_param { top_k: 1 } '
        'include { phase: TEST } }')
    o.w('layer { name: "top5" type: "Accuracy" bottom: "fc8" bottom: "label" top: "top5" accuracy_param { top_k: 5 } '
        'include { phase: TEST } }')
    return o.f.getvalue()


def _full_body(o, data_blob, segments, classes, fc_name, dropout3d, dropout2d):
    b3 = _trunk_to_3c(o, data_blob)
    # inception_3c is split around the 3-D head exactly as in the reference file
    t = _cbr2d(o, "inception_3c", "3x3_reduce", b3, 128, 1)
    c3 = _cbr2d(o, "inception_3c", "3x3", t, 160, 3, 2, 1)
    t = _cbr2d(o, "inception_3c", "double_3x3_reduce", b3, 64, 1)
    d1 = _cbr2d(o, "inception_3c", "double_3x3_1", t, 96, 3, None, 1)
    _head3d(o, d1, segments)
    d2 = _cbr2d(o, "inception_3c", "double_3x3_2", d1, 96, 3, 2, 1)
    _pool(o, "inception_3c_pool", b3, "MAX", 3, 2)
    _concat(o, "inception_3c_output", [c3, d2, "inception_3c_pool"])
    t = "inception_3c_output"
    t = _inception(o, "4a", t, 224, 64, 96, 96, 128, 128, "AVE", 128)
    t = _inception(o, "4b", t, 192, 96, 128, 96, 128, 128, "AVE", 128)
    t = _inception(o, "4c", t, 160, 128, 160, 128, 160, 160, "AVE", 128)
    t = _inception(o, "4d", t, 96, 128, 192, 160, 192, 192, "AVE", 128)
    t = _inception(o, "4e", t, 0, 128, 192, 192, 256, 256, "MAX", 0, stride2=True)
    t = _inception(o, "5a", t, 352, 192, 320, 160, 224, 224, "AVE", 128)
    t = _inception(o, "5b", t, 352, 192, 320, 192, 224, 224, "MAX", 128)
    _pool(o, "global_pool2D", t, "AVE", 7, 1)
    o.w('layer { name: "dropout2D" type: "Dropout" bottom: "global_pool2D" top: "global_pool2D"')
    o.w("  dropout_param { dropout_ratio: %g } }" % dropout2d)
    o.w('layer { name: "reshape_fc_st2" type: "Reshape" bottom: "global_pool2D" top: "reshape_fc_st2"')
    o.w("  reshape_param { shape { dim: -1 dim: 1 dim: %d dim: 1024 } } }" % segments)
    o.w('layer { name: "segment_consensus_st2" type: "Pooling" bottom: "reshape_fc_st2" top: "pool_fusion_st2"')
    o.w("  pooling_param { pool: AVE kernel_h: %d kernel_w: 1 } }" % segments)
    o.w('layer { name: "global_pool_reshape2D" type: "Reshape" bottom: "pool_fusion_st2" top: "pool_fusion_st2D"')
    o.w("  reshape_param { shape { dim: -1 dim: 1024 } } }")
    _tail3d(o, segments, dropout3d)
    _concat(o, "gn02_concat", ["pool_fusion_st2D", "global_pool_reshape"], top="global_pool_gn02_reshape", axis=1)
    _fc(o, fc_name, "global_pool_gn02_reshape", classes)


def _train_header(o, net_name, batch, segments):
    o.w('name: "%s"' % net_name)
    o.w('input: "data"')
    for d in (batch, 3 * segments, 224, 224):
        o.w("input_dim: %d" % d)
    o.w('input: "label"')
    for d in (batch, 1, 1, 1):
        o.w("input_dim: %d" % d)
    o.w('layer { name: "reshape_data" type: "Reshape" bottom: "data" top: "reshape_data" '
        'reshape_param { shape { dim: -1 dim: 3 dim: 224 dim: 224 } } }')


def _loss_tail(o):
    o.w('layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc8" bottom: "label" include { phase: TRAIN } top: "loss" }')
    o.w('layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc8" bottom: "label" top: "loss" include { phase: TEST } }')
    o.w('layer { name: "top1" type: "Accuracy" bottom: "fc8" bottom: "label" top: "top1" accuracy_param { top_k: 1 } '
        'include { phase: TEST } }')
    o.w('layer { name: "top5" type: "Accuracy" bottom: "fc8" bottom: "label" top: "top5" accuracy_param { top_k: 5 } '
        'include { phase: TEST } }')


def eco_full_deploy(segments=16, classes=400, fc_name="fc8N", batch=5, dropout3d=0.5, dropout2d=0.6,
                    net_name="o3d"):
    """ECO-Full deploy net (cf. models_ECO_Full/kinetics/deploy.prototxt): the Lite graph plus
    the 2-D stream inception_3c..5b -> global_pool2D -> segment consensus, concatenated
    with the 3-D feature before the classifier."""
    o = _W()
    _header(o, net_name, batch * segments)
    _full_body(o, "data", segments, classes, fc_name, dropout3d, dropout2d)
    return o.f.getvalue()


def eco_full_train(segments=16, classes=400, fc_name="fc8N", batch=8, dropout3d=0.5, dropout2d=0.6, net_name="o3d"):
    """ECO-Full train/test net (cf. models_ECO_Full/kinetics/ECO_full.prototxt) with the VideoData layers replaced by the two
    net inputs they produce (see eco_lite_train)."""
    o = _W()
    _train_header(o, net_name, batch, segments)
    _full_body(o, "reshape_data", segments, classes, fc_name, dropout3d, dropout2d)
    _loss_tail(o)
    return o.f.getvalue()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", choices=["lite", "full"])
    ap.add_argument("--segments", type=int, default=16)
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--fc", default=None)
    ap.add_argument("--batch", type=int, default=5, help="videos in the declared input_dim")
    ap.add_argument("--dropout", type=float, default=None)
    a = ap.parse_args()
    if a.model == "lite":
        print(eco_lite_deploy(a.segments, a.classes or 101, a.fc or "fc8u", a.batch,
                              0.6 if a.dropout is None else a.dropout), end="")
    else:
        print(eco_full_deploy(a.segments, a.classes or 400, a.fc or "fc8N", a.batch,
                              0.5 if a.dropout is None else a.dropout), end="")


if __name__ == "__main__":
    main()
