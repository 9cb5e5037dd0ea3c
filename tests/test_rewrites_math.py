"""CPU checks of the algebra behind the device plan's rewrites, done with the oracle's own operators
(oracle/ref_cpu.c through oracle.refnet), independent of any GPU:

* the 7x7/stride-2/pad-3 stem == a 4x4/stride-1 convolution over 2x2 space-to-depth cells (what `stem_rows_kernel`
  computes), with exactly the weight packing `Net::upload_params` uses, and == the sum over the 4 vertical taps of
  "cell row y+i times weight K block i" (the sliding-row accumulation into the TMEM ring);
* `pool_commute`: AVE 3x3/s1/p1 pooling followed by a 1x1 convolution == the 1x1 convolution without bias followed by
  the pooling plus bias (caffe's AVE divides every window by 9 here and padding contributes zeros);
* `fuse_1x1`: sibling 1x1 convolutions as one GEMM over concatenated output channels;
* training: the input gradient of a strided 1x1 convolution == W^T dY on the compact output grid scattered to every S-th
  input position (net_train.inc: `compact`); the input gradient of a strided 3x3x3 convolution == the stride-1 convolution of
  the zero-dilated dY with the flipped, transposed filter (`setup_dgrad`); the bias gradient of a convolution read only by a
  batch-statistics BN is zero up to rounding (`db_zero`).
"""
import numpy as np

from oracle import refnet


def cells_of(x):
    """[F,3,H,W] -> space-to-depth cells [F, CH, CW, 16] as csrc/aux_kernels.cu:stem_s2d_kernel lays them out:
    cell (Y,X), value (dy*2+dx)*3+c = x[c, 2Y+dy-3, 2X+dx-3], zero outside the frame, values 12..15 zero."""
    F, C, H, W = x.shape
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    CH, CW = OH + 3, OW + 3
    xp = np.zeros((F, C, 2 * CH + 3, 2 * CW + 3), np.float32)
    xp[:, :, 3:3 + H, 3:3 + W] = x  # xp[..., y+3, x+3] = x[..., y, x]
    cells = np.zeros((F, CH, CW, 16), np.float32)
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                cells[..., (dy * 2 + dx) * 3 + c] = xp[:, c, dy:dy + 2 * CH:2, dx:dx + 2 * CW:2]
    return cells, OH, OW


def pack_stem_weights(w):
    """[Cout,3,7,7] -> [Cout, 4 vertical taps, 64] with K index tx*16 + (dy*2+dx)*3 + c  (net.cpp: upload_params)"""
    cout = w.shape[0]
    wp = np.zeros((cout, 4, 64), np.float32)
    for ty in range(4):
        for tx in range(4):
            for dy in range(2):
                for dx in range(2):
                    ky, kx = 2 * ty + dy, 2 * tx + dx
                    if ky < 7 and kx < 7:
                        for c in range(3):
                            wp[:, ty, tx * 16 + (dy * 2 + dx) * 3 + c] = w[:, c, ky, kx]
    return wp


def test_stem_is_a_sliding_window_over_cell_rows():
    rng = np.random.default_rng(0)
    for hw in ((32, 32), (30, 34), (17, 23)):
        x = rng.normal(size=(2, 3) + hw).astype(np.float32)
        w = rng.normal(size=(8, 3, 7, 7)).astype(np.float32)
        b = rng.normal(size=(8,)).astype(np.float32)
        want = refnet.conv_forward(x, w, b, [7, 7], [2, 2], [3, 3])
        cells, OH, OW = cells_of(x)
        wp = pack_stem_weights(w)
        got = np.zeros((2, 8, OH, OW), np.float32)
        for y in range(OH):           # output row y = sum over the 4 cell rows y..y+3 (vertical taps)
            for i in range(4):
                row = cells[:, y + i]  # [F, CW, 16]
                # window x = cells x..x+3 of this row = 64 consecutive values (what one 128-byte tile row holds)
                win = np.stack([row[:, xx:xx + 4].reshape(2, 64) for xx in range(OW)], axis=1)  # [F, OW, 64]
                got[:, :, y, :] += np.einsum("fxk,ok->fox", win, wp[:, i])
        got += b[None, :, None, None]
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max(), hw


def test_ave_pool_then_1x1_conv_commutes():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(3, 24, 9, 11)).astype(np.float32)
    w = rng.normal(size=(16, 24, 1, 1)).astype(np.float32)
    b = rng.normal(size=(16,)).astype(np.float32)
    pooled = refnet.pool_forward(x, [3, 3], [1, 1], [1, 1], "AVE")
    want = refnet.conv_forward(pooled, w, b, [1, 1], [1, 1], [0, 0])
    nobias = refnet.conv_forward(x, w, np.zeros_like(b), [1, 1], [1, 1], [0, 0])
    got = refnet.pool_forward(nobias, [3, 3], [1, 1], [1, 1], "AVE") + b[None, :, None, None]
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # with a bias inside the conv the two orders differ at the border (padding is zero, not bias): the rewrite must
    # keep the bias behind the pooling
    wrong = refnet.pool_forward(refnet.conv_forward(x, w, b, [1, 1], [1, 1], [0, 0]), [3, 3], [1, 1], [1, 1], "AVE")
    assert np.abs(wrong - want).max() > 1e-2 * np.abs(want).max()


def test_sibling_1x1_convs_as_one_gemm():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2, 32, 6, 7)).astype(np.float32)
    ws = [rng.normal(size=(n, 32, 1, 1)).astype(np.float32) for n in (16, 32, 16)]
    bs = [rng.normal(size=(w.shape[0],)).astype(np.float32) for w in ws]
    fused = refnet.conv_forward(x, np.concatenate(ws), np.concatenate(bs), [1, 1], [1, 1], [0, 0])
    off = 0
    for w, b in zip(ws, bs):
        one = refnet.conv_forward(x, w, b, [1, 1], [1, 1], [0, 0])
        assert np.array_equal(fused[:, off:off + w.shape[0]], one)  # same K order per output channel: bit-identical
        off += w.shape[0]


def test_strided_1x1_input_gradient_is_a_scatter_of_the_compact_gemm():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 24, 5, 9, 8)).astype(np.float32)
    w = rng.normal(size=(40, 24, 1, 1, 1)).astype(np.float32)
    k, s, p = [1, 1, 1], [2, 2, 2], [0, 0, 0]
    od, oh, ow = (5 - 1) // 2 + 1, (9 - 1) // 2 + 1, (8 - 1) // 2 + 1
    dy = rng.normal(size=(2, 40, od, oh, ow)).astype(np.float32)
    dx, _, _ = refnet.conv_backward(x, w, dy, k, s, p)
    compact = np.einsum("oc,nozyx->nczyx", w[:, :, 0, 0, 0], dy)      # W^T dY on the output grid
    want = np.zeros_like(x)
    want[:, :, ::2, ::2, ::2][:, :, :od, :oh, :ow] = compact
    assert np.abs(dx - want).max() <= 1e-4 * np.abs(want).max()
    # positions the strided convolution never read get exactly zero
    mask = np.ones(x.shape[2:], bool)
    mask[::2, ::2, ::2] = False
    assert not np.any(dx[:, :, mask])


def test_strided_input_gradient_is_a_stride1_conv_of_the_dilated_gradient():
    rng = np.random.default_rng(6)
    cin, cout, I = 6, 10, (4, 7, 6)
    x = rng.normal(size=(1, cin) + I).astype(np.float32)
    w = rng.normal(size=(cout, cin, 3, 3, 3)).astype(np.float32)
    k, s, p = [3, 3, 3], [2, 2, 2], [1, 1, 1]
    O = tuple((i + 2 - 3) // 2 + 1 for i in I)
    dy = rng.normal(size=(1, cout) + O).astype(np.float32)
    dx, _, _ = refnet.conv_backward(x, w, dy, k, s, p)
    E = tuple(i + 2 - 3 + 1 for i in I)                                 # dilated extent: I + 2 pad - K + 1
    dil = np.zeros((1, cout) + E, np.float32)
    dil[:, :, ::2, ::2, ::2][:, :, :O[0], :O[1], :O[2]] = dy
    wt = np.ascontiguousarray(np.flip(w, axis=(2, 3, 4)).transpose(1, 0, 2, 3, 4))  # [cin][cout][flipped taps]
    got = refnet.conv_forward(dil, wt, np.zeros(cin, np.float32), k, [1, 1, 1], [1, 1, 1])   # pad_new = K - 1 - pad
    assert got.shape == dx.shape
    assert np.abs(got - dx).max() <= 1e-4 * np.abs(dx).max()


def test_bias_gradient_in_front_of_a_batch_statistics_bn_is_rounding_noise():
    rng = np.random.default_rng(7)
    x = rng.normal(size=(4, 16, 6, 6)).astype(np.float32) * 3 + 1
    slope = rng.uniform(0.5, 1.5, 16).astype(np.float32)
    dy = rng.normal(size=x.shape).astype(np.float32)
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    dx, _, _ = refnet.bn_backward_train(x, dy, slope, mean, var)
    colsum = dx.sum(axis=(0, 2, 3))                                      # what backward_cpu_bias would accumulate
    assert np.abs(colsum).max() <= 1e-4 * np.abs(dx).sum(axis=(0, 2, 3)).max()


def test_strided_3x3_input_gradient_as_parity_class_convolutions():
    """dX of a 3x3x3 / stride 2 / pad 1 convolution split by the parity of the input position -- class
    (pz,py,px) is a STRIDE-1 convolution of the compact dY with the 1- or 2-tap sub-filter per axis that has the matching
    parity (27 taps in total over the 8 classes), instead of 27 taps over a dilated map that is 7/8 zeros."""
    rng = np.random.default_rng(8)
    cin, cout, I = 5, 7, (4, 6, 8)
    x = rng.normal(size=(1, cin) + I).astype(np.float32)
    w = rng.normal(size=(cout, cin, 3, 3, 3)).astype(np.float32)
    O = tuple((i + 2 - 3) // 2 + 1 for i in I)
    dy = rng.normal(size=(1, cout) + O).astype(np.float32)
    want, _, _ = refnet.conv_backward(x, w, dy, [3, 3, 3], [2, 2, 2], [1, 1, 1])
    got = np.zeros_like(want)
    taps_total = 0
    # input index i = 2a + q reads dY[(i + 1 - k) / 2] for the k with i + 1 - k even: q = 0 -> k = 1 (dY[a]);
    # q = 1 -> k = 0 (dY[a + 1]) and k = 2 (dY[a])
    sub = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}   # parity -> [(tap k, offset into the compact grid)]
    dyp = np.pad(dy, ((0, 0), (0, 0), (0, 1), (0, 1), (0, 1)))   # a + 1 may run one past the grid: zero
    for qz in (0, 1):
        for qy in (0, 1):
            for qx in (0, 1):
                nz, ny, nx = len(range(qz, I[0], 2)), len(range(qy, I[1], 2)), len(range(qx, I[2], 2))
                acc = np.zeros((1, cin, nz, ny, nx), np.float32)
                for kz, oz in sub[qz]:
                    for ky, oy in sub[qy]:
                        for kx, ox in sub[qx]:
                            taps_total += 1
                            g = dyp[:, :, oz:oz + nz, oy:oy + ny, ox:ox + nx]
                            acc += np.einsum("oc,nozyx->nczyx", w[:, :, kz, ky, kx], g)
                got[:, :, qz::2, qy::2, qx::2] = acc
    assert taps_total == 27
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
