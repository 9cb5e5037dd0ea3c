"""Online sliding-window recognition with a per-frame feature cache (BASELINE config #5).

The reference's demo (scripts/online_recognition/online_recognition.py:64-93) keeps the last N frames and calls
net.forward() on all of them for every new frame.  Through the 2-D trunk frames are independent (TEST-phase BN), so only
the NEW frames need the trunk: their inception_3c_double_3x3_1_bn features (96 x 28 x 28 bf16 = 150 KB per frame) are
appended to a window that lives in the full net's own feature blob on the device, and only the 3-D head (r2Dto3D ...
fc) runs per step.  Per step: 1.8 GFLOP x new frames + 64 GFLOP (N=16) instead of 93 GFLOP.  Results are bit-identical
to the uncached forward on the same window (tests/test_gpu_online.py)."""
from __future__ import annotations

import re

import numpy as np

from ._caffe import lib
from .pycaffe import Net, TEST

FEATURE_BLOB = "inception_3c_double_3x3_1_bn"


def _with_frames(prototxt_text, frames):
    """the deploy definition with its first input_dim (frames = videos x segments) replaced"""
    out, n = re.subn(r"(input_dim:\s*)\d+", lambda m: m.group(1) + str(frames), prototxt_text, count=1)
    if n == 0:
        out, n = re.subn(r"(input_shape\s*\{\s*dim:\s*)\d+", lambda m: m.group(1) + str(frames), prototxt_text, count=1)
    if n == 0:
        raise ValueError("prototxt has no input_dim / input_shape to resize")
    return out


class SlidingWindowRecognizer(object):
    def __init__(self, prototxt_text, segments, new_frames=1, weights=None, feature_blob=FEATURE_BLOB, output="fc8", **options):
        self.segments, self.k, self.feature_blob, self.output = int(segments), int(new_frames), feature_blob, output
        self.full = Net.from_string(_with_frames(prototxt_text, self.segments), TEST, **options)
        self.trunk = Net.from_string(_with_frames(prototxt_text, self.k), TEST, until_blob=feature_blob, **options)
        if weights:
            self.full.copy_from(weights)
        self.trunk.share_with(self.full)
        L = lib()
        fb = self.full._blob_names.index(feature_blob)
        last_writer = max(i for i in range(len(self.full._layer_names))
                          if fb in [L.eco_net_layer_top(self.full._h, i, j) for j in range(L.eco_net_layer_num_tops(self.full._h, i))])
        self.head_start = last_writer + 1
        self.filled = 0
        self.input = self.trunk.inputs[0]

    def load_params_from(self, net):
        self.full.share_with(net)
        self.trunk.share_with(net)

    def push(self, frames):
        """frames: [k, 3, H, W] fp32 (mean-subtracted, as the data layer hands over).  Returns the logits of the current window
        once N frames have been seen, else None."""
        frames = np.asarray(frames, np.float32)
        assert frames.shape[0] == self.k, (frames.shape, self.k)
        self.trunk.blobs[self.input].data[...] = frames
        self.trunk._forward(0, len(self.trunk.layers) - 1)
        self.full.push_frames(self.feature_blob, self.trunk, self.feature_blob)
        self.filled = min(self.segments, self.filled + self.k)
        if self.filled < self.segments:
            return None
        self.full._forward(self.head_start, len(self.full.layers) - 1)
        return self.full.blobs[self.output].data

    def forward_window_uncached(self, window):
        """what the reference does: the whole net on all N frames"""
        self.full.blobs[self.full.inputs[0]].data[...] = window
        return self.full.forward()[self.output]
