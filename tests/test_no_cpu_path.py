"""The product has no CPU execution path: without a CUDA device the graph can be built and inspected (prototxt
parsing, shape inference, parameter access are host work) but `forward` must fail loudly -- never fall back to the
oracle or any other CPU implementation (the oracle is test infrastructure only)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "eco-efficient-video-understanding_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

NET = """name: "t"
input: "data" input_dim: 1 input_dim: 8 input_dim: 6 input_dim: 6
layer { name: "c" type: "Convolution" bottom: "data" top: "c" convolution_param { num_output: 16 kernel_size: 3 pad: 1 } }
layer { name: "c_bn" type: "BN" bottom: "c" top: "c_bn" }
layer { name: "c_relu" type: "ReLU" bottom: "c_bn" top: "c_bn" }
"""


def test_forward_without_cuda_raises_and_never_computes():
    import caffe
    if caffe.device_count() > 0:
        pytest.skip("a CUDA device is visible: this test covers the device-less behaviour")
    net = caffe.Net.from_string(NET, caffe.TEST)
    # host-side work is available without a device
    assert list(net.blobs) == ["data", "c", "c_bn"]
    assert tuple(net.blobs["c_bn"].shape) == (1, 16, 6, 6)
    assert tuple(net.params["c"][0].shape) == (16, 8, 3, 3)
    net.blobs["data"].data[...] = np.ones((1, 8, 6, 6), np.float32)
    with pytest.raises(RuntimeError) as e:
        net.forward()
    assert "CUDA" in str(e.value) or "cuda" in str(e.value) or "device" in str(e.value)


def test_product_sources_do_not_reach_into_the_oracle():
    # only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use oracle/
    pkg = os.path.join(ROOT, "eco-efficient-video-understanding_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base or os.sep + "lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if "oracle" in txt and ("import oracle" in txt or "from oracle" in txt or "ref_cpu" in txt or "libref_cpu" in txt):
                    offenders.append(os.path.join(base, f))
    for f in ("include/eco_b200.h", "include/eco_caffe_facade.hpp"):
        txt = open(os.path.join(ROOT, f)).read()
        if "libref_cpu" in txt or "ref_cpu" in txt:
            offenders.append(f)
    assert not offenders, offenders
