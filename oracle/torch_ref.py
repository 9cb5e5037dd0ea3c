"""oracle/torch_ref.py -- TEST INFRASTRUCTURE.  ECO-Lite written directly with torch's CPU fp32 functional ops
(oneDNN / MKL): (1) the second, independent pin of the oracle (tests/test_oracle_vs_torch.py; SURVEY.md 8(c):
whole-network outputs are not pinned by any reference fixture, so two independent restatements must agree), and
(2) the "strong CPU" line SURVEY.md 8(d) asks for next to the caffe-algorithm port (bench.py cpu_baseline leg).
Layer semantics follow the same reference files as oracle/ref_cpu.c (conv_layer.cpp:12-43, bn_layer.cpp:93-170,
pooling_layer.cpp:131-147,195-262 -> ceil_mode / count_include_pad)."""
import numpy as np
import torch
import torch.nn.functional as F


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def torch_eco_lite(params, x, segments):
    """params: {layer: [arrays]} (RefNet.params_dict() layout); x: torch tensor [B*N,3,224,224]."""
    P = {n: [t(a) for a in arrs] for n, arrs in params.items()}

    def cbr(x, name, stride=1, pad=0, nd=2):
        fn = F.conv2d if nd == 2 else F.conv3d
        y = fn(x, P[name][0], P[name][1], stride=stride, padding=pad)
        return bnrelu(y, name + "_bn")

    def bnrelu(y, bn):
        g, b, m, v = [p.reshape([1, -1] + [1] * (y.dim() - 2)) for p in P[bn]]
        return F.relu((y - m) * (v + 1e-5).pow(-0.5) * g + b)

    h = cbr(x, "conv1_7x7_s2", 2, 3)
    h = F.max_pool2d(h, 3, 2, ceil_mode=True)
    h = cbr(h, "conv2_3x3_reduce")
    h = cbr(h, "conv2_3x3", 1, 1)
    h = F.max_pool2d(h, 3, 2, ceil_mode=True)
    for blk in ("3a", "3b"):
        p = "inception_" + blk
        b1 = cbr(h, p + "_1x1")
        b2 = cbr(cbr(h, p + "_3x3_reduce"), p + "_3x3", 1, 1)
        b3 = cbr(cbr(cbr(h, p + "_double_3x3_reduce"), p + "_double_3x3_1", 1, 1), p + "_double_3x3_2", 1, 1)
        b4 = cbr(F.avg_pool2d(h, 3, 1, 1, count_include_pad=True), p + "_pool_proj")
        h = torch.cat([b1, b2, b3, b4], 1)
    h = cbr(cbr(h, "inception_3c_double_3x3_reduce"), "inception_3c_double_3x3_1", 1, 1)
    h = h.reshape(-1, segments, 96, 28, 28).permute(0, 2, 1, 3, 4).contiguous()
    conv3 = lambda x, n, s: F.conv3d(x, P[n][0], P[n][1], stride=s, padding=1)
    res3a = conv3(h, "res3a_2n", 1)
    u = conv3(bnrelu(conv3(bnrelu(res3a, "res3a_bn"), "res3b_1", 1), "res3b_1_bn"), "res3b_2", 1)
    tcur = bnrelu(u + res3a, "res3b_bn")
    for st in ("res4", "res5"):
        a, b = st + "a", st + "b"
        ra = conv3(bnrelu(conv3(tcur, a + "_1", 2), a + "_1_bn"), a + "_2", 1) + conv3(tcur, a + "_down", 2)
        ta = bnrelu(ra, a + "_bn")
        rb = conv3(bnrelu(conv3(ta, b + "_1", 1), b + "_1_bn"), b + "_2", 1) + ra
        tcur = bnrelu(rb, b + "_bn")
    feat = tcur.mean((2, 3, 4))
    fc = [n for n in P if n.startswith("fc8")][0]
    return (feat @ P[fc][0].t() + P[fc][1]).numpy(), tcur.numpy()
