"""The step in front of the hot path: what the reference's VideoData layer + DataTransformer do on the host
(caffe_3d/src/caffe/layers/video_data_layer.cpp:134-238, data_transformer.cpp:50-326), with the pixel work on the GPU.

    feeder = VideoFeeder(net, 'data', 'label', segments=16, train=True, seed=1,
                         transform=dict(mirror=1, multi_scale=1, fix_crop=1, more_fix_crop=1, max_distort=1,
                                        scale_ratios=[1, .875, .75, .66], mean_value=[104, 117, 123]))
    offsets = feeder.sample_offsets(num_frames)          # which frames of a video to decode (TSN segment sampling)
    feeder.feed(clips_u8, labels)                        # clips: uint8 [B, 3*segments, H, W] (decoded frames, Datum layout)

JPEG decoding / list shuffling / prefetch threads stay with the caller; crop (multi-scale sizes x fixed offsets), bilinear
resize (OpenCV's 8-bit fixed-point INTER_LINEAR), mirror and mean subtraction run in one kernel that writes the net's
fp32 input blob on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._caffe import check, lib


class ClipTransform(C.Structure):
    _fields_ = [("h_off", C.c_int), ("w_off", C.c_int), ("crop_h", C.c_int), ("crop_w", C.c_int), ("mirror", C.c_int)]


class TransformParam(C.Structure):
    _fields_ = [("mirror", C.c_int), ("multi_scale", C.c_int), ("fix_crop", C.c_int), ("more_fix_crop", C.c_int),
                ("max_distort", C.c_int), ("is_flow", C.c_int), ("num_scale_ratios", C.c_int), ("scale_ratios", C.c_float * 8),
                ("scale", C.c_float), ("num_mean", C.c_int), ("mean_value", C.c_float * 16)]

    @classmethod
    def make(cls, mirror=0, multi_scale=0, fix_crop=0, more_fix_crop=0, max_distort=1, is_flow=0, scale_ratios=(), scale=1.0,
             mean_value=()):
        p = cls()
        p.mirror, p.multi_scale, p.fix_crop, p.more_fix_crop = int(mirror), int(multi_scale), int(fix_crop), int(more_fix_crop)
        p.max_distort, p.is_flow, p.scale = int(max_distort), int(is_flow), float(scale)
        p.num_scale_ratios = len(scale_ratios)
        for i, r in enumerate(scale_ratios):
            p.scale_ratios[i] = float(r)
        mv = list(mean_value)[:16]
        p.num_mean = len(mv)
        for i, m in enumerate(mv):
            p.mean_value[i] = float(m)
        return p


def _bind():
    L = lib()
    if getattr(L, "_video_bound", False):
        return L
    L.eco_sampler_create.argtypes = [C.c_uint, C.POINTER(C.c_void_p)]
    L.eco_sampler_destroy.argtypes = [C.c_void_p]
    L.eco_sample_segment_offsets.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.eco_sample_clip_transform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TransformParam), C.POINTER(ClipTransform)]
    L.eco_crop_size_candidates.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.eco_fix_offset_candidates.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.eco_net_transform_input_u8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(ClipTransform), C.POINTER(TransformParam)]
    L._video_bound = True
    return L


def crop_size_candidates(H, W, crop_size, max_distort=1, ratios=()):
    L = _bind()
    buf = (C.c_int * 128)()
    n = C.c_int(64)
    r = (C.c_float * max(1, len(ratios)))(*[float(x) for x in ratios])
    check(L.eco_crop_size_candidates(H, W, crop_size, max_distort, r if ratios else None, len(ratios), buf, C.byref(n)))
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]


def fix_offset_candidates(H, W, crop_h, crop_w, more):
    L = _bind()
    buf = (C.c_int * 64)()
    n = C.c_int(32)
    check(L.eco_fix_offset_candidates(H, W, crop_h, crop_w, int(more), buf, C.byref(n)))
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]


class Sampler(object):
    """the two mt19937 streams of the data layer: frame sampling and transform choices (`rng() % n`, like caffe::rng_t)"""

    def __init__(self, seed):
        self._h = C.c_void_p()
        check(_bind().eco_sampler_create(int(seed), C.byref(self._h)))

    def __del__(self):
        try:
            if self._h.value:
                lib().eco_sampler_destroy(self._h)
        except Exception:
            pass

    def segment_offsets(self, num_frames, num_segments, new_length=1, train=True):
        out = (C.c_int * num_segments)()
        check(_bind().eco_sample_segment_offsets(self._h, num_frames, num_segments, new_length, int(train), out))
        return list(out)

    def clip_transform(self, H, W, crop_size, train, param):
        t = ClipTransform()
        check(_bind().eco_sample_clip_transform(self._h, H, W, crop_size, int(train), C.byref(param), C.byref(t)))
        return t


def transform_into(net, blob, clips_u8, transforms, param):
    """clips_u8: uint8 [B, C, H, W]; transforms: list of ClipTransform; writes net.blobs[blob] on the device"""
    clips_u8 = np.ascontiguousarray(clips_u8, np.uint8)
    B, Cc, H, W = clips_u8.shape
    arr = (ClipTransform * B)(*transforms)
    check(_bind().eco_net_transform_input_u8(net._h, net._blob_names.index(blob), clips_u8.ctypes.data_as(C.c_void_p), B, Cc, H, W,
                                             arr, C.byref(param)))


class VideoFeeder(object):
    def __init__(self, net, data_blob="data", label_blob="label", segments=16, train=True, seed=1, transform=None):
        self.net, self.data_blob, self.label_blob = net, data_blob, label_blob
        self.segments, self.train = segments, train
        self.param = TransformParam.make(**(transform or {}))
        self.sampler = Sampler(seed)
        self.crop = net.blobs[data_blob].shape[2]

    def sample_offsets(self, num_frames, new_length=1):
        return self.sampler.segment_offsets(num_frames, self.segments, new_length, self.train)

    def feed(self, clips_u8, labels=None):
        B, _, H, W = clips_u8.shape
        ts = [self.sampler.clip_transform(H, W, self.crop, self.train, self.param) for _ in range(B)]
        transform_into(self.net, self.data_blob, clips_u8, ts, self.param)
        if labels is not None and self.label_blob in self.net.blobs:
            self.net.blobs[self.label_blob].data[...] = np.asarray(labels, np.float32).reshape(self.net.blobs[self.label_blob].shape)
        return ts
