"""DataTransformer::Transform on the GPU against the oracle's restatement, bit-exact: crop windows of every
multi-scale size, fixed offsets, mirror, flow inversion, mean replication, scale; then the feeder in front of a real net."""
import numpy as np
import pytest

from oracle import transform_ref as T

pytestmark = pytest.mark.gpu

NET = 'name: "t"\ninput: "data"\ninput_dim: %d\ninput_dim: %d\ninput_dim: %d\ninput_dim: %d\n' \
      'layer { name: "c" type: "Convolution" bottom: "data" top: "c" convolution_param { num_output: 8 kernel_size: 1 } }\n'


@pytest.mark.parametrize("is_flow", [0, 1])
def test_transform_kernel_bit_exact(gpu, is_flow):
    import caffe
    from caffe.video_data import ClipTransform, TransformParam, transform_into
    B, C, H, W, crop = 5, 6, 64, 85, 56
    net = caffe.Net.from_string(NET % (B, C, crop, crop), caffe.TEST)
    rng = np.random.default_rng(1)
    clips = rng.integers(0, 256, (B, C, H, W), dtype=np.uint8)
    cands = T.crop_size_candidates(H, W, crop, crop, 1)
    ts, want = [], []
    for b in range(B):
        ch, cw = cands[(3 * b + 1) % len(cands)]
        ho, wo = T.fix_offset_candidates(H, W, ch, cw, True)[(5 * b + 2) % 13]
        t = dict(h_off=ho, w_off=wo, crop_h=ch, crop_w=cw, mirror=b % 2)
        ts.append(ClipTransform(ho, wo, ch, cw, b % 2))
        want.append(T.transform(clips[b], t, crop, mean_values=(104, 117, 123), scale=0.5, is_flow=bool(is_flow)))
    p = TransformParam.make(mirror=1, multi_scale=1, fix_crop=1, more_fix_crop=1, is_flow=is_flow, scale=0.5, mean_value=[104, 117, 123])
    transform_into(net, "data", clips, ts, p)
    got = np.array(net.blobs["data"].data)
    assert np.array_equal(got, np.stack(want)), "max diff %g" % np.abs(got - np.stack(want)).max()


def test_feeder_drives_a_net(gpu):
    import caffe
    from caffe.video_data import VideoFeeder
    B, seg, H, W, crop = 2, 4, 64, 85, 56
    net = caffe.Net.from_string(NET % (B, 3 * seg, crop, crop), caffe.TEST)
    feeder = VideoFeeder(net, "data", None, segments=seg, train=True, seed=11,
                         transform=dict(mirror=1, multi_scale=1, fix_crop=1, more_fix_crop=1, mean_value=[104, 117, 123]))
    assert len(feeder.sample_offsets(97)) == seg
    clips = np.random.default_rng(2).integers(0, 256, (B, 3 * seg, H, W), dtype=np.uint8)
    ts = feeder.feed(clips)
    want = np.stack([T.transform(clips[b], dict(h_off=t.h_off, w_off=t.w_off, crop_h=t.crop_h, crop_w=t.crop_w, mirror=t.mirror), crop,
                                 mean_values=(104, 117, 123)) for b, t in enumerate(ts)])
    out = net.forward()["c"]            # the transformed blob is what the net reads: no host upload overwrites it
    assert np.array_equal(np.array(net.blobs["data"].data), want)
    assert np.isfinite(out).all()
