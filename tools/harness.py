"""Harness-owned random weights for benchmarks and demos, generated WITHOUT the oracle: only numpy and the
product's own public surface (caffe.Net.params / blobs / forward).

Distributions follow SURVEY.md 8(d): conv/fc W ~ U(+-sqrt(3/fan_in)) (xavier, filler.hpp:149-163), bias ~ N(0, 0.01),
BN slope ~ U(0.5, 1.5), bias ~ N(0, 0.1).  caffe's own fillers leave the BN running statistics at 0/0 (bn_layer.cpp:34-41),
which overflows after a few layers; a trained net has statistics that match its activations, so the harness
sets every BN layer's running mean / variance from the activations the DEVICE produces on one synthetic clip."""
from __future__ import annotations

import numpy as np


def synthetic_frames(batch_videos, segments, seed=1234, size=224):
    """uint8 U{0..255} frames, fp32, BGR mean (104,117,123) subtracted, no scale: what the data layer hands over."""
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, size=(batch_videos * segments, 3, size, size), dtype=np.uint8)
    return u8.astype(np.float32) - np.array([104.0, 117.0, 123.0], np.float32).reshape(1, 3, 1, 1)


def init_params(net, seed=4321):
    rng = np.random.default_rng(seed)
    types = {name: lr.type for name, lr in zip(net._layer_names, net.layers)}
    for name, blobs in net.params.items():
        t = types[name]
        if t in ("Convolution", "InnerProduct"):
            shp = tuple(blobs[0].shape)
            a = np.sqrt(3.0 / int(np.prod(shp[1:])))
            blobs[0].data[...] = rng.uniform(-a, a, shp).astype(np.float32)
            if len(blobs) > 1:
                blobs[1].data[...] = rng.normal(0, 0.01, tuple(blobs[1].shape)).astype(np.float32)
        elif t == "BN":
            blobs[0].data[...] = rng.uniform(0.5, 1.5, tuple(blobs[0].shape)).astype(np.float32)
            blobs[1].data[...] = rng.normal(0, 0.1, tuple(blobs[1].shape)).astype(np.float32)
            blobs[2].data[...] = 0.0
            blobs[3].data[...] = 1.0
    return net


def calibrate_bn_on_device(net, x, seed=99):
    """`net` must materialise every blob (option keep_all_blobs=1).  Walks the BN layers in order; each gets
    the per-channel mean / variance of its input as the device computes it (with a small jitter so that the
    normalised activations are not exactly zero-mean / unit-variance)."""
    from caffe import _caffe
    import ctypes as C
    rng = np.random.default_rng(seed)
    L = _caffe.lib()
    blobs = net.blobs
    net.blobs[net.inputs[0]].data[...] = x
    for li, (name, lr) in enumerate(zip(net._layer_names, net.layers)):
        if lr.type != "BN":
            continue
        net.blobs[net.inputs[0]].data[...] = x
        net.forward()
        b = blobs[net._blob_names[L.eco_net_layer_bottom(net._h, li, 0)]].data
        ch = b.shape[1]
        xs = np.moveaxis(b, 1, 0).reshape(ch, -1).astype(np.float64)
        m, v = xs.mean(1), xs.var(1)
        m = m + rng.normal(0, 0.1, ch) * np.sqrt(v + 1e-5)
        v = np.maximum(v * rng.uniform(0.8, 1.25, ch), 1e-4)
        p = net.params[name]
        p[2].data[...] = m.astype(np.float32).reshape(p[2].shape)
        p[3].data[...] = v.astype(np.float32).reshape(p[3].shape)
    return net


def copy_params(dst, src):
    """same architecture, any batch size / segment count: parameter shapes do not depend on them"""
    sp = src.params
    for name, blobs in dst.params.items():
        for a, b in zip(blobs, sp[name]):
            a.data[...] = b.data
    return dst


def params_dict(net):
    return {name: [np.array(b.data, np.float32, copy=True) for b in blobs] for name, blobs in net.params.items()}
