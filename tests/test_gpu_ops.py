"""GPU parity tests for the individual fused ops, driven through the public surface
(caffe shim -> C ABI -> sm_100a kernels) on tiny nets, checked against the oracle
(oracle/ref_cpu.c) and against the reference's golden pooling vectors.
Every conv case runs in both A-operand modes (cp.async gather / TMA im2col)."""
import json
import os

import numpy as np
import pytest

from oracle import refnet
from eco_testlib import check_bf16_blob, check_f32_blob, load_params, make_net, teacher_blobs, teacher_raw_blobs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def conv_net(shape, cout, k, s, p, bn=True, relu=True):
    nsp = len(shape) - 2
    dims = "".join("input_dim: %d\n" % d for d in shape) if len(shape) == 4 else \
        "input_shape { %s }\n" % " ".join("dim: %d" % d for d in shape)
    lst = lambda v: "[%s]" % ", ".join(str(i) for i in v)
    txt = 'name: "t"\ninput: "data"\n' + dims
    txt += ('layer { name: "c" type: "Convolution" bottom: "data" top: "c" convolution_param { num_output: %d '
            'kernel_size: %s stride: %s pad: %s } }\n' % (cout, lst(k), lst(s), lst(p)))
    if bn:
        txt += 'layer { name: "c_bn" type: "BN" bottom: "c" top: "c_bn" }\n'
        if relu:
            txt += 'layer { name: "c_relu" type: "ReLU" bottom: "c_bn" top: "c_bn" }\n'
    return txt


def two_conv_net(shape, cmid, cout, k, s, p):
    """fp32 input -> 1x1 conv (+BN+ReLU) -> conv under test (+BN+ReLU): the second conv reads a
    channels-last bf16 tensor exactly like every conv inside ECO."""
    nsp = len(shape) - 2
    one = [1] * nsp
    zero = [0] * nsp
    txt = conv_net(shape, cmid, one, one, zero).replace('"c"', '"a"').replace('"c_bn"', '"a_bn"').replace('"c_relu"', '"a_relu"')
    lst = lambda v: "[%s]" % ", ".join(str(i) for i in v)
    txt += ('layer { name: "c" type: "Convolution" bottom: "a_bn" top: "c" convolution_param { num_output: %d '
            'kernel_size: %s stride: %s pad: %s } }\n' % (cout, lst(k), lst(s), lst(p)))
    txt += 'layer { name: "c_bn" type: "BN" bottom: "c" top: "c_bn" }\n'
    txt += 'layer { name: "c_relu" type: "ReLU" bottom: "c_bn" top: "c_bn" }\n'
    return txt


PLAIN = {"gp", "gp_r", "fc"}  # blobs that are plain fp32 on the device


def run_case(txt, shape, a_mode, check=("c_bn",), seed=0, keep_all=True, persistent=True, dual_m=1, halo=1,
             stem_rows=None, **extra):
    ref = refnet.RefNet(txt).init_params(seed + 1)
    rng = np.random.default_rng(seed)
    x = rng.normal(size=shape).astype(np.float32)
    want = ref.forward(x, bf16=True)
    net = make_net(txt, keep_all=keep_all, a_mode=a_mode, persistent=persistent, dual_m=dual_m, halo=halo,
                   stem_rows=stem_rows, **extra)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    net.forward()
    got = net.blobs
    dev = {name: got[name].data.copy() for name in check}
    # teacher forcing: every checked blob is recomputed by the oracle from the device's own upstream
    # blobs, so each fused op is judged on identical inputs
    forced = ref.forward(x, bf16=True, teacher=teacher_blobs(ref, dev), teacher_raw=teacher_raw_blobs(ref, dev))
    for name in check:
        g = dev[name]
        assert g.shape == want[name].shape, (name, g.shape, want[name].shape)
        if name in PLAIN:
            check_f32_blob(g, forced[name], name)
        else:
            check_bf16_blob(g, forced[name], name)
    return net, want


CONV2D = [
    # shape, cmid, cout, k, s, p
    ((2, 16, 12, 12), 64, 64, [1, 1], [1, 1], [0, 0]),      # pure GEMM, K=64, M=288 (ragged last tile)
    ((2, 16, 12, 12), 64, 32, [3, 3], [1, 1], [1, 1]),      # 3x3 p1
    ((1, 8, 28, 28), 192, 64, [1, 1], [1, 1], [0, 0]),      # inception_3a_1x1 geometry (3 K blocks)
    ((2, 8, 14, 14), 96, 96, [3, 3], [1, 1], [1, 1]),       # Cin=96: half-empty second K block (zero fill)
    ((2, 8, 15, 15), 64, 160, [3, 3], [2, 2], [1, 1]),      # stride 2, odd size, N=160
    ((1, 8, 9, 9), 64, 352, [1, 1], [1, 1], [0, 0]),        # Cout 352 -> two N tiles of 176
    ((1, 8, 7, 7), 320, 192, [1, 1], [1, 1], [0, 0]),       # K=320 (5 blocks), N=192
]


@pytest.mark.parametrize("mode", ["gather", "im2col", "halo"])
@pytest.mark.parametrize("case", CONV2D, ids=["c%d" % i for i in range(len(CONV2D))])
def test_conv2d(gpu, case, mode):
    # "im2col": per-tap TMA im2col loads; "halo": stride-1 filters read a patch loaded once (auto-selected
    # when eligible, otherwise identical to im2col)
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 0 if mode == "gather" else 1, check=("a_bn", "c", "c_bn"),
             halo=1 if mode == "halo" else 0)


HALO = [
    ((2, 8, 28, 28), 64, 64, [3, 3], [1, 1], [1, 1]),       # inception_3a_3x3
    ((2, 8, 28, 28), 96, 96, [3, 3], [1, 1], [1, 1]),       # double_3x3_2: Cin 96 (zero-filled second block)
    ((1, 8, 56, 56), 64, 192, [3, 3], [1, 1], [1, 1]),      # conv2_3x3
    ((3, 8, 14, 14), 128, 160, [3, 3], [1, 1], [1, 1]),     # ECO-Full 4c geometry
    ((2, 8, 20, 23), 64, 32, [5, 3], [1, 1], [2, 1]),       # rectangular filter, ragged last band
    ((2, 8, 17, 9), 64, 64, [3, 3], [1, 1], [0, 0]),        # no padding
    ((1, 8, 12, 12), 256, 352, [3, 3], [1, 1], [1, 1]),     # four channel blocks, two N tiles
]


@pytest.mark.parametrize("mode", ["auto", "mt2", "stream"])
@pytest.mark.parametrize("case", HALO, ids=["h%d" % i for i in range(len(HALO))])
def test_conv2d_halo_kernel(gpu, case, mode):
    # auto: weights resident in smem when they fit (else the im2col kernel); mt2: two 128-position halves per
    # tile forced; stream: the halo kernel even when the weights must be streamed per tap
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"),
             halo={"auto": 1, "mt2": 2, "stream": 3}[mode])


CONV3D = [
    ((1, 8, 4, 10, 10), 64, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1]),
    ((2, 8, 8, 14, 14), 128, 256, [3, 3, 3], [2, 2, 2], [1, 1, 1]),   # res4a_1 geometry (stride 2)
    ((1, 8, 4, 7, 7), 96, 128, [3, 3, 3], [1, 1, 1], [1, 1, 1]),      # res3a_2n: Cin 96
    ((3, 8, 2, 7, 7), 256, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1]),     # res5 geometry: M=294, two N tiles
]


@pytest.mark.parametrize("a_mode", [0, 1], ids=["gather", "im2col"])
@pytest.mark.parametrize("case", CONV3D, ids=["d%d" % i for i in range(len(CONV3D))])
def test_conv3d(gpu, case, a_mode):
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, a_mode, check=("a_bn", "c", "c_bn"))


@pytest.mark.parametrize("case", [CONV2D[1], CONV2D[5]], ids=["c1", "c5"])
def test_conv2d_im2col_one_tile_per_cta(gpu, case):
    # the non-persistent kernel with the TMA im2col A path (the persistent one is the default)
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"), persistent=False)


@pytest.mark.parametrize("case", [CONV2D[0], CONV2D[1], CONV2D[3]], ids=["c0", "c1", "c3"])
def test_conv2d_dual_m_tiles(gpu, case):
    # 256-row tiles: two 128-row halves share each weight tile (forced; auto needs >= 2 tiles per SM)
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"), dual_m=2)


@pytest.mark.parametrize("case", [CONV2D[0], CONV2D[1], CONV2D[3], CONV2D[5]], ids=["c0", "c1", "c3", "c5"])
def test_conv2d_cta_pair_kernel(gpu, case):
    # cta_group::2: two CTAs of a cluster share one 256-row tile, each loads half of the weight tile (forced; the
    # cost model only picks it for layers with at least a tile per SM pair)
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"), pair=2)


@pytest.mark.parametrize("case", CONV3D, ids=["d%d" % i for i in range(len(CONV3D))])
def test_conv3d_cta_pair_kernel(gpu, case):
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"), pair=2)


def test_cta_pair_residual_and_many_tiles(gpu):
    run_case(RES_NET, (2, 8, 4, 6, 6), 1, check=("ra", "ra_bn", "rb_bn", "rc_bn", "fc"), pair=2)
    shape = (8, 8, 8, 14, 14)  # more 256-row tiles than SM pairs, two N tiles: both TMEM buffers, ring wrap
    run_case(two_conv_net(shape, 128, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1]), shape, 1, check=("c", "c_bn"), pair=2)


@pytest.mark.parametrize("dual_m", [1, 2], ids=["mt1", "mt2"])
@pytest.mark.parametrize("case", [CONV2D[0], CONV2D[1], CONV2D[5], CONV3D[0], CONV3D[2], CONV3D[3]],
                         ids=["c0", "c1", "c5", "d0", "d2", "d3"])
def test_conv_multicast_clusters(gpu, case, dual_m):
    # persistent kernel in clusters of two CTAs: different M tiles of the same N tile, each CTA loads half of the
    # weight tile and multicasts it into both shared memories (forced; auto needs >= 2 tiles per SM).  Odd tile
    # counts leave one CTA of the last cluster with a dead tile that still feeds its weight half to the peer.
    shape, cmid, cout, k, s, p = case
    run_case(two_conv_net(shape, cmid, cout, k, s, p), shape, 1, check=("a_bn", "c", "c_bn"), multicast=2, dual_m=dual_m, pair=0)


def test_multicast_clusters_residual_and_many_tiles(gpu):
    run_case(RES_NET, (2, 8, 4, 6, 6), 1, check=("ra", "ra_bn", "rb_bn", "rc_bn", "fc"), multicast=2, pair=0)
    shape = (8, 8, 8, 14, 14)
    run_case(two_conv_net(shape, 128, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1]), shape, 1, check=("c", "c_bn"), multicast=2, pair=0)
    run_case(INCEPTION, (2, 8, 14, 14), 1, check=("b1_bn", "b2_bn", "nx_bn"), keep_all=False, multicast=2, pool_commute=0)


def test_conv3d_dual_m_residual(gpu):
    run_case(RES_NET, (2, 8, 4, 6, 6), 1, check=("ra", "ra_bn", "rb_bn", "rc_bn", "fc"), dual_m=2)


def test_conv3d_many_tiles_persistent(gpu):
    # more tiles than SMs and two N tiles: every CTA of the persistent kernel walks several tiles,
    # both TMEM accumulator buffers and the per-tile constant reload are exercised
    shape = (8, 8, 8, 14, 14)
    run_case(two_conv_net(shape, 128, 512, [3, 3, 3], [1, 1, 1], [1, 1, 1]), shape, 1, check=("c", "c_bn"))


@pytest.mark.parametrize("mode", ["gather", "im2col", "halo", "halo_mt2"])
@pytest.mark.parametrize("hw", [(32, 32), (30, 34), (224, 224)])
def test_stem_7x7_s2(gpu, hw, mode):
    # conv1_7x7_s2 on fp32 NCHW input: space-to-depth transform, then either the 4x1 overlapping-window view
    # (gather / im2col) or the halo kernel over 16-channel cells (32-byte swizzled rows, resident weights)
    shape = (2, 3) + hw
    run_case(conv_net(shape, 64, [7, 7], [2, 2], [3, 3]), shape, 0 if mode == "gather" else 1, check=("c", "c_bn"),
             halo={"gather": 0, "im2col": 0, "halo": 1, "halo_mt2": 2}[mode])


STEM_POOL = 'layer { name: "p1" type: "Pooling" bottom: "c_bn" top: "p1" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }\n'


@pytest.mark.parametrize("shape", [(2, 3, 32, 32), (2, 3, 30, 34), (3, 3, 224, 224), (160, 3, 32, 32), (1, 3, 256, 250)],
                         ids=["32x32", "30x34", "224", "160frames", "256x250"])
@pytest.mark.parametrize("direct", [1, 0], ids=["frames", "cells"])
@pytest.mark.parametrize("pool", [False, True], ids=["rows", "rows_pool"])
def test_stem_rows_kernel(gpu, shape, pool, direct):
    # the production stem: sliding window over cell rows with a ring of TMEM accumulators; with `pool` the MAX
    # 3x3/s2 pooling behind it (pool1) is applied on chip and the full-resolution map is never stored.
    # 160 frames: more work units than SMs (several units per CTA, both rings wrap); odd / even output heights
    # cover the clipped last pooling window of caffe's ceil mode.
    txt = conv_net(shape, 64, [7, 7], [2, 2], [3, 3]) + (STEM_POOL if pool else "")
    # frames: the kernel bulk-copies raw image rows and builds the windows on chip (widths that are a multiple of 16;
    # other widths and `stem_direct=0` go through the cell buffer written by the transform kernel + tiled TMA)
    net, _ = run_case(txt, shape, 1, check=("p1",) if pool else ("c_bn",), keep_all=False, stem_rows=1, stem_direct=direct)
    with pytest.raises(RuntimeError):
        net.blobs["c"].data  # never materialised by this kernel
    if pool:
        with pytest.raises(RuntimeError):
            net.blobs["c_bn"].data


@pytest.mark.parametrize("warps", [5, 6])
def test_stem_rows_kernel_gather_warps(gpu, warps):
    # the window-gather role of the direct stem kernel with 5 / 6 warps (default 4): rows are dealt round-robin
    shape = (5, 3, 224, 224)
    txt = conv_net(shape, 64, [7, 7], [2, 2], [3, 3]) + STEM_POOL
    run_case(txt, shape, 1, check=("p1",), keep_all=False, stem_rows=1, stem_direct=1, stem_gather_warps=warps)


def test_stem_rows_pool_not_folded_when_blob_has_other_readers(gpu):
    # c_bn feeds the pooling AND a 1x1 conv: the pool must stay a separate op (c_bn is stored)
    shape = (2, 3, 32, 32)
    txt = conv_net(shape, 64, [7, 7], [2, 2], [3, 3]) + STEM_POOL
    txt += ('layer { name: "d" type: "Convolution" bottom: "c_bn" top: "d" convolution_param { num_output: 32 '
            'kernel_size: 1 } }\n')
    run_case(txt, shape, 1, check=("c_bn", "p1", "d"), keep_all=False, stem_rows=1)


@pytest.mark.parametrize("a_mode", [0, 1], ids=["gather", "im2col"])
def test_generic_fp32_input_conv(gpu, a_mode):
    # a first conv that is not the 7x7 stem: channels padded to 8 on the fly
    shape = (2, 5, 9, 11)
    run_case(conv_net(shape, 24, [3, 3], [1, 1], [1, 1]), shape, a_mode, check=("c", "c_bn"))


def test_conv_without_bn_and_plain_relu(gpu):
    shape = (1, 8, 6, 6)
    txt = conv_net(shape, 16, [3, 3], [1, 1], [1, 1], bn=False)
    run_case(txt, shape, 0, check=("c",))
    txt2 = txt + 'layer { name: "r" type: "ReLU" bottom: "c" top: "c" }\n'
    run_case(txt2, shape, 0, check=("c",))


def test_bn_after_in_place_relu_is_not_fused_into_the_conv(gpu):
    """conv -> ReLU (in place on the conv top) -> BN: the reference computes BN(relu(conv)).  The planner may fuse a BN
    that reads the conv top into the conv epilogue only when no in-place layer rewrites that top in between
    (round-1 advisor finding: it used to produce relu-after-BN silently)."""
    shape = (2, 8, 6, 6)
    txt = conv_net(shape, 32, [3, 3], [1, 1], [1, 1], bn=False)
    txt += 'layer { name: "r" type: "ReLU" bottom: "c" top: "c" }\n'
    txt += 'layer { name: "c_bn" type: "BN" bottom: "c" top: "c_bn" }\n'
    for keep_all in (True, False):
        run_case(txt, shape, 0, check=("c_bn",) if not keep_all else ("c", "c_bn"), keep_all=keep_all)


RES_NET = """name: "res"
input: "data" input_shape { dim: 2 dim: 8 dim: 4 dim: 6 dim: 6 }
layer { name: "s" type: "Convolution" bottom: "data" top: "s" convolution_param { num_output: 64 kernel_size: [1,1,1] } }
layer { name: "s_bn" type: "BN" bottom: "s" top: "s_bn" }
layer { name: "s_relu" type: "ReLU" bottom: "s_bn" top: "s_bn" }
layer { name: "ra_2n" type: "Convolution" bottom: "s_bn" top: "ra" convolution_param { num_output: 64 kernel_size: [3,3,3] pad: [1,1,1] } }
layer { name: "ra_bn" type: "BN" bottom: "ra" top: "ra_bn" }
layer { name: "ra_relu" type: "ReLU" bottom: "ra_bn" top: "ra_bn" }
layer { name: "rb_1" type: "Convolution" bottom: "ra_bn" top: "rb_1" convolution_param { num_output: 64 kernel_size: [3,3,3] pad: [1,1,1] } }
layer { name: "rb_1_bn" type: "BN" bottom: "rb_1" top: "rb_1_bn" }
layer { name: "rb_1_relu" type: "ReLU" bottom: "rb_1_bn" top: "rb_1_bn" }
layer { name: "rb_2" type: "Convolution" bottom: "rb_1_bn" top: "rb_2" convolution_param { num_output: 64 kernel_size: [3,3,3] pad: [1,1,1] } }
layer { name: "rb" type: "Eltwise" bottom: "rb_2" bottom: "ra" top: "rb" }
layer { name: "rb_bn" type: "BN" bottom: "rb" top: "rb_bn" }
layer { name: "rb_relu" type: "ReLU" bottom: "rb_bn" top: "rb_bn" }
layer { name: "rc_1" type: "Convolution" bottom: "rb_bn" top: "rc_1" convolution_param { num_output: 128 kernel_size: [3,3,3] pad: [1,1,1] stride: [2,2,2] } }
layer { name: "rc_1_bn" type: "BN" bottom: "rc_1" top: "rc_1_bn" }
layer { name: "rc_1_relu" type: "ReLU" bottom: "rc_1_bn" top: "rc_1_bn" }
layer { name: "rc_2" type: "Convolution" bottom: "rc_1_bn" top: "rc_2" convolution_param { num_output: 128 kernel_size: [3,3,3] pad: [1,1,1] } }
layer { name: "rc_down" type: "Convolution" bottom: "rb_bn" top: "rc_down" convolution_param { num_output: 128 kernel_size: [3,3,3] pad: [1,1,1] stride: [2,2,2] } }
layer { name: "rc" type: "Eltwise" bottom: "rc_2" bottom: "rc_down" top: "rc" }
layer { name: "rc_bn" type: "BN" bottom: "rc" top: "rc_bn" }
layer { name: "rc_relu" type: "ReLU" bottom: "rc_bn" top: "rc_bn" }
layer { name: "gp" type: "Pooling" bottom: "rc_bn" top: "gp" pooling_param { pool: AVE kernel_size: [2,3,3] stride: [1,1,1] } }
layer { name: "gp_r" type: "Reshape" bottom: "gp" top: "gp_r" reshape_param { shape { dim: -1 dim: 128 } } }
layer { name: "drop" type: "Dropout" bottom: "gp_r" top: "gp_r" dropout_param { dropout_ratio: 0.5 } }
layer { name: "fc" type: "InnerProduct" bottom: "gp_r" top: "fc" inner_product_param { num_output: 10 } }
"""


@pytest.mark.parametrize("keep_all", [True, False])
@pytest.mark.parametrize("a_mode", [0, 1], ids=["gather", "im2col"])
def test_residual_block_fusion(gpu, a_mode, keep_all):
    # pre-activation residual structure of the 3-D head: raw sums feed the shortcut,
    # BN+ReLU after each add; both fused into the conv epilogues.
    shape = (2, 8, 4, 6, 6)
    check = ("ra", "ra_bn", "rb_bn", "rc_2", "rc_bn", "gp", "fc") if keep_all else ("fc",)
    run_case(RES_NET, shape, a_mode, check=check, keep_all=keep_all)


INCEPTION = """name: "inc"
input: "data" input_dim: 2 input_dim: 8 input_dim: 14 input_dim: 14
layer { name: "stem" type: "Convolution" bottom: "data" top: "stem" convolution_param { num_output: 64 kernel_size: 1 } }
layer { name: "stem_bn" type: "BN" bottom: "stem" top: "stem_bn" }
layer { name: "stem_relu" type: "ReLU" bottom: "stem_bn" top: "stem_bn" }
layer { name: "b1" type: "Convolution" bottom: "stem_bn" top: "b1" convolution_param { num_output: 32 kernel_size: 1 } }
layer { name: "b1_bn" type: "BN" bottom: "b1" top: "b1_bn" }
layer { name: "b1_relu" type: "ReLU" bottom: "b1_bn" top: "b1_bn" }
layer { name: "b2r" type: "Convolution" bottom: "stem_bn" top: "b2r" convolution_param { num_output: 64 kernel_size: 1 } }
layer { name: "b2r_bn" type: "BN" bottom: "b2r" top: "b2r_bn" }
layer { name: "b2r_relu" type: "ReLU" bottom: "b2r_bn" top: "b2r_bn" }
layer { name: "b2" type: "Convolution" bottom: "b2r_bn" top: "b2" convolution_param { num_output: 96 kernel_size: 3 pad: 1 } }
layer { name: "b2_bn" type: "BN" bottom: "b2" top: "b2_bn" }
layer { name: "b2_relu" type: "ReLU" bottom: "b2_bn" top: "b2_bn" }
layer { name: "pl" type: "Pooling" bottom: "stem_bn" top: "pl" pooling_param { pool: AVE kernel_size: 3 stride: 1 pad: 1 } }
layer { name: "pp" type: "Convolution" bottom: "pl" top: "pp" convolution_param { num_output: 32 kernel_size: 1 } }
layer { name: "pp_bn" type: "BN" bottom: "pp" top: "pp_bn" }
layer { name: "pp_relu" type: "ReLU" bottom: "pp_bn" top: "pp_bn" }
layer { name: "mp" type: "Pooling" bottom: "stem_bn" top: "mp" pooling_param { pool: MAX kernel_size: 3 stride: 1 pad: 1 } }
layer { name: "out" type: "Concat" bottom: "b1_bn" bottom: "b2_bn" bottom: "pp_bn" bottom: "mp" top: "out" }
layer { name: "down" type: "Pooling" bottom: "out" top: "down" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }
layer { name: "nx" type: "Convolution" bottom: "down" top: "nx" convolution_param { num_output: 64 kernel_size: 1 } }
layer { name: "nx_bn" type: "BN" bottom: "nx" top: "nx_bn" }
"""


@pytest.mark.parametrize("a_mode", [0, 1], ids=["gather", "im2col"])
def test_inception_block_concat_alias(gpu, a_mode):
    # concat is zero-copy: branch epilogues / pools write channel slices of `out`
    run_case(INCEPTION, (2, 8, 14, 14), a_mode, check=("b1_bn", "b2_bn", "pl", "pp_bn", "mp", "out", "down", "nx_bn"))


@pytest.mark.parametrize("shape", [(2, 8, 14, 14), (5, 8, 28, 28), (3, 8, 9, 11)], ids=["14", "28", "9x11"])
def test_inception_pool_conv_commuted_in_fast_plan(gpu, shape):
    # fast plan: AVE 3x3/s1/p1 pooling -> 1x1 conv -> BN -> ReLU runs as conv (64 -> 32 channels, no bias) -> pooling
    # with bias + BN + ReLU in its epilogue (exact by linearity; one bf16 rounding moves from the pooled blob to the
    # conv output).  The pooled blob is never formed; the concat slice must match the oracle to bf16 noise.
    from eco_testlib import rel_l2
    txt = INCEPTION.replace("input_dim: 2 input_dim: 8 input_dim: 14 input_dim: 14",
                            " ".join("input_dim: %d" % d for d in shape))
    ref = refnet.RefNet(txt).init_params(5)
    x = np.random.default_rng(3).normal(size=shape).astype(np.float32)
    want = ref.forward(x, bf16=True)
    got = {}
    for commute in (1, 0):
        net = make_net(txt, keep_all=False, a_mode=1)
        net.set_option("pool_commute", commute)
        load_params(net, ref.params_dict())
        net.blobs["data"].data[...] = x
        net.forward()
        got[commute] = {k: net.blobs[k].data.copy() for k in ("pp_bn", "out", "nx_bn")}
        if commute:
            with pytest.raises(RuntimeError):
                net.blobs["pl"].data  # never formed
        else:
            assert net.blobs["pl"].data.shape == want["pl"].shape
    for k in ("pp_bn", "out", "nx_bn"):
        e1, e0 = rel_l2(got[1][k], want[k]), rel_l2(got[0][k], want[k])
        assert e1 <= 6e-3 and e0 <= 6e-3, (k, e1, e0)
        scale = np.abs(want[k]).max()
        assert np.abs(got[1][k] - want[k]).max() <= 2e-2 * scale, k
    # sibling 1x1 fusion (b1 + b2r + the commuted pp run as one GEMM with three output segments) changes no bit
    net = make_net(txt, keep_all=False, a_mode=1)
    net.set_option("fuse_1x1", 0)
    load_params(net, ref.params_dict())
    net.blobs["data"].data[...] = x
    net.forward()
    for k in ("pp_bn", "out", "nx_bn"):
        assert np.array_equal(net.blobs[k].data, got[1][k]), k
    # the other branches are untouched by the rewrite: bit-identical between the two plans
    c32 = 32 + 96
    assert np.array_equal(got[1]["out"][:, :c32], got[0]["out"][:, :c32])
    assert np.array_equal(got[1]["out"][:, c32 + 32:], got[0]["out"][:, c32 + 32:])


POOLS = """name: "pools"
input: "data" input_dim: 3 input_dim: 8 input_dim: %d input_dim: %d
layer { name: "f" type: "Convolution" bottom: "data" top: "f" convolution_param { num_output: 24 kernel_size: 1 } }
layer { name: "f_bn" type: "BN" bottom: "f" top: "f_bn" }
layer { name: "p_max_s2" type: "Pooling" bottom: "f_bn" top: "p_max_s2" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }
layer { name: "p_max_s2p" type: "Pooling" bottom: "f_bn" top: "p_max_s2p" pooling_param { pool: MAX kernel_size: 3 stride: 2 pad: 1 } }
layer { name: "p_max_s1p" type: "Pooling" bottom: "f_bn" top: "p_max_s1p" pooling_param { pool: MAX kernel_size: 3 stride: 1 pad: 1 } }
layer { name: "p_ave_s1p" type: "Pooling" bottom: "f_bn" top: "p_ave_s1p" pooling_param { pool: AVE kernel_size: 3 stride: 1 pad: 1 } }
layer { name: "p_ave_s2p" type: "Pooling" bottom: "f_bn" top: "p_ave_s2p" pooling_param { pool: AVE kernel_size: 3 stride: 2 pad: 1 } }
layer { name: "p_ave_s2" type: "Pooling" bottom: "f_bn" top: "p_ave_s2" pooling_param { pool: AVE kernel_size: 3 stride: 2 } }
layer { name: "p_ave_5" type: "Pooling" bottom: "f_bn" top: "p_ave_5" pooling_param { pool: AVE kernel_size: 5 stride: 3 pad: 2 } }
layer { name: "p_max_hw" type: "Pooling" bottom: "f_bn" top: "p_max_hw" pooling_param { pool: MAX kernel_h: 2 kernel_w: 3 stride: 1 } }
"""


@pytest.mark.parametrize("hw", [(15, 13), (28, 28), (7, 9), (112, 112)])
def test_pool_variants(gpu, hw):
    # the 3x3 strip kernel (stride 1/2, pad 0/1, ragged strips, ceil-mode last window) and the generic kernel
    shape = (3, 8) + hw
    run_case(POOLS % hw, shape, 1, check=("f_bn", "p_max_s2", "p_max_s2p", "p_max_s1p", "p_ave_s1p", "p_ave_s2p", "p_ave_s2",
                                          "p_ave_5", "p_max_hw"))


GOLD = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_pooling_reference_known_answers_on_device(gpu, case):
    # the reference's golden pooling vectors through the channels-last device kernel:
    # an identity 1x1 conv lifts the fp32 input into a bf16 channels-last blob (small integers are exact)
    nsp = len(case["in_shape"])
    shape = [2, 8] + case["in_shape"]
    lst = lambda v: "[%s]" % ", ".join(str(i) for i in v)
    txt = 'name: "p"\ninput: "data"\ninput_shape { %s }\n' % " ".join("dim: %d" % d for d in shape)
    txt += 'layer { name: "id" type: "Convolution" bottom: "data" top: "id" convolution_param { num_output: 8 kernel_size: %s bias_term: false } }\n' % lst([1] * nsp)
    txt += ('layer { name: "pool" type: "Pooling" bottom: "id" top: "pool" pooling_param { pool: %s kernel_size: %s '
            'stride: %s pad: %s } }\n' % (case["method"], lst(case["kernel"]), lst(case["stride"]), lst(case["pad"])))
    net = make_net(txt, keep_all=True, a_mode=0)
    net.params["id"][0].data[...] = np.eye(8, dtype=np.float32).reshape([8, 8] + [1] * nsp)
    x1 = np.array(case["input"], np.float32).reshape(case["in_shape"])
    net.blobs["data"].data[...] = np.broadcast_to(x1, shape)
    net.forward()
    got = net.blobs["pool"].data
    assert list(got.shape) == [2, 8] + case["out_shape"]
    exp = np.array(case["output"], np.float32).reshape(case["out_shape"])
    tol = max(case["tol"], 8e-3 * np.abs(exp).max())  # bf16 storage of the averaged values (rel 2^-9)
    if case["method"] == "MAX":
        tol = case["tol"]
    assert np.abs(got - exp).max() <= tol + 1e-12
