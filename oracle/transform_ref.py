"""oracle/transform_ref.py -- TEST INFRASTRUCTURE.  CPU restatement of the step in front of the hot path:

  * VideoDataLayer's segment sampling          caffe_3d/src/caffe/layers/video_data_layer.cpp:155-187
  * fillCropSize / fillFixOffset               caffe_3d/src/caffe/data_transformer.cpp:83-105, :50-78
  * DataTransformer::Transform (Datum)         data_transformer.cpp:148-326: crop window -> cv::resize to crop_size (per channel
                                               plane, only when the window is not already crop_size) -> mirror -> flow inversion
                                               -> (v - mean[c]) * scale

cv::resize is a third-party dependency of the reference (OpenCV, version unpinned by the reference's build files): its
INTER_LINEAR 8-bit path is restated here from OpenCV's published algorithm (imgproc/resize.cpp: 11-bit fixed-point
coefficients, HResizeLinear then VResizeLinear with ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2) and pinned against the
real cv2.resize of this image (OpenCV 4.13) by tests/test_transform_cpu.py and the committed tests/golden/resize_cv2.npz."""
import numpy as np


def _coef(dsize, ssize):
    scale = np.float32(ssize) / np.float32(dsize)
    d = np.arange(dsize, dtype=np.float32)
    fx = (d + np.float32(0.5)) * scale - np.float32(0.5)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0
    fx[lo] = 0
    sx[lo] = 0
    hi = sx >= ssize - 1
    fx[hi] = 0
    sx[hi] = ssize - 1
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    return sx, np.minimum(sx + 1, ssize - 1), a0, a1


def resize_linear_u8(plane, dh, dw):
    """cv::resize(src, dst, Size(dw, dh)) with INTER_LINEAR for an 8-bit single-channel image"""
    plane = np.asarray(plane, np.uint8)
    sh, sw = plane.shape
    x0, x1, ax0, ax1 = _coef(dw, sw)
    y0, y1, by0, by1 = _coef(dh, sh)
    p = plane.astype(np.int64)
    H = p[:, x0] * ax0[None, :] + p[:, x1] * ax1[None, :]        # horizontal pass, 8 + 11 bits
    S0, S1 = H[y0, :], H[y1, :]
    v = (((by0[:, None] * (S0 >> 4)) >> 16) + ((by1[:, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def crop_size_candidates(H, W, net_h, net_w, max_distort=1, ratios=(1.0, .875, .75, .66)):
    out = []
    base = min(H, W)
    for i, rh in enumerate(ratios):
        ch = int(base * np.float32(rh))
        ch = net_h if abs(ch - net_h) < 3 else ch
        for j, rw in enumerate(ratios):
            cw = int(base * np.float32(rw))
            cw = net_w if abs(cw - net_w) < 3 else cw
            if abs(i - j) <= max_distort:
                out.append((ch, cw))
    return out


def fix_offset_candidates(H, W, crop_h, crop_w, more):
    ho, wo = (H - crop_h) // 4, (W - crop_w) // 4
    o = [(0, 0), (0, 4 * wo), (4 * ho, 0), (4 * ho, 4 * wo), (2 * ho, 2 * wo)]
    if more:
        o += [(0, 2 * wo), (4 * ho, 2 * wo), (2 * ho, 0), (2 * ho, 4 * wo), (ho, wo), (ho, 3 * wo), (3 * ho, wo), (3 * ho, 3 * wo)]
    return o


class MT19937(object):
    """std::mt19937 / boost::mt19937 (caffe::rng_t) stream: numpy's legacy RandomState seeds the same generator"""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def __call__(self):
        return int(self.rs._bit_generator.random_raw())


def segment_offsets(num_frames, num_segments, new_length, train, rng):
    avg = float(num_frames // num_segments)
    out = []
    for i in range(num_segments):
        if train:
            if avg >= new_length:
                off = rng() % (int(avg) - new_length + 1)
                out.append(int(off + i * avg))
            else:
                out.append(int(i * avg))
        else:
            out.append(int((avg - new_length + 1) / 2 + i * avg) if avg >= new_length else 0)
    return out


def sample_transform(H, W, crop_size, train, rng, mirror=True, multi_scale=True, fix_crop=True, more_fix_crop=True, max_distort=1,
                     ratios=(1.0, .875, .75, .66)):
    t = {"mirror": 1 if (mirror and rng() % 2) else 0}
    if train:
        if multi_scale:
            cs = crop_size_candidates(H, W, crop_size, crop_size, max_distort, ratios)
            t["crop_h"], t["crop_w"] = cs[rng() % len(cs)]
        else:
            t["crop_h"] = t["crop_w"] = crop_size
        if fix_crop:
            os_ = fix_offset_candidates(H, W, t["crop_h"], t["crop_w"], more_fix_crop)
            t["h_off"], t["w_off"] = os_[rng() % len(os_)]
        else:
            t["h_off"] = rng() % (H - t["crop_h"] + 1)
            t["w_off"] = rng() % (W - t["crop_w"] + 1)
    else:
        t["crop_h"] = t["crop_w"] = crop_size
        t["h_off"], t["w_off"] = (H - crop_size) // 2, (W - crop_size) // 2
    return t


def transform(datum, t, crop_size, mean_values=(), scale=1.0, is_flow=False):
    """datum: uint8 [C, H, W]; t: dict(h_off, w_off, crop_h, crop_w, mirror) -> fp32 [C, crop, crop]"""
    datum = np.asarray(datum, np.uint8)
    C = datum.shape[0]
    out = np.empty((C, crop_size, crop_size), np.float32)
    mv = list(mean_values)
    if len(mv) == 1:
        mv = mv * C
    elif mv and len(mv) < C:
        mv = [mv[c % len(mean_values)] for c in range(C)]
    for c in range(C):
        win = datum[c, t["h_off"]:t["h_off"] + t["crop_h"], t["w_off"]:t["w_off"] + t["crop_w"]]
        if (t["crop_h"], t["crop_w"]) != (crop_size, crop_size):
            win = resize_linear_u8(win, crop_size, crop_size)
        e = win.astype(np.float32)
        if is_flow and t["mirror"] and c < C // 2:
            e = np.float32(255) - e
        if t["mirror"]:
            e = e[:, ::-1]
        m = np.float32(mv[c]) if mv else np.float32(0)
        out[c] = (e - m) * np.float32(scale)
    return out
