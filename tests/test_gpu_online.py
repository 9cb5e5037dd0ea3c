"""Online sliding window with the per-frame feature cache (BASELINE config #5): bit-identical logits to the
uncached forward of the same window, for N = 8 (README recipe) and N = 16 (what the reference script runs), one or two
new frames per step; and the uncached path itself against the oracle."""
import numpy as np
import pytest

from oracle import refnet
import gen_eco_prototxt as gen
from eco_testlib import TOL_LOGITS, load_params, rel_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("segments,k", [(8, 1), (16, 1), (16, 2)])
def test_cached_sliding_window_equals_full_forward(gpu, segments, k):
    from caffe.online import SlidingWindowRecognizer
    txt = gen.eco_lite_deploy(segments=segments, classes=101, batch=1)
    ref = refnet.RefNet(txt).init_params(4321)
    stream = refnet.eco_input(1, segments + 3 * k + k, seed=77)          # a video longer than the window
    ref.calibrate_bn(stream[:segments])
    rec = SlidingWindowRecognizer(txt, segments, new_frames=k)
    load_params(rec.full, ref.params_dict())
    rec.trunk.share_with(rec.full)
    outs = []
    for t in range(0, stream.shape[0] - k + 1, k):
        r = rec.push(stream[t:t + k])
        if r is not None:
            outs.append((t + k, r.copy()))
    assert len(outs) >= 3
    plain = __import__("caffe").Net.from_string(txt, 1)
    load_params(plain, ref.params_dict())
    for end, got in outs:
        window = stream[end - segments:end]
        plain.blobs["data"].data[...] = window
        want = plain.forward()["fc8"].copy()
        assert np.array_equal(got, want), "window ending at frame %d: max diff %g" % (end, np.abs(got - want).max())
    # and that forward is the oracle's (bf16 mirror), for the last window
    end, got = outs[-1]
    w = ref.forward(stream[end - segments:end], bf16=True)["fc8"]
    assert rel_max(got, w) <= TOL_LOGITS
