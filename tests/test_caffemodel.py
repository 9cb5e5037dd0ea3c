"""caffemodel I/O (csrc/caffemodel.cpp, no libprotobuf) -- round trip through the C ABI on the CPU, and,
where /root/reference is mounted, against the reference's own generated schema
(caffe_3d/python/caffe/proto/caffe_pb2.py) in both directions: what Net::ToProto / CopyTrainedLayersFrom
(caffe_3d/src/caffe/net.cpp:852-904) would write and read."""
import os
import subprocess
import sys

import numpy as np
import pytest

import caffe
import gen_eco_prototxt as gen
from oracle import refnet

REF_PROTO = "/root/reference/caffe_3d/python/caffe/proto"


def small_net():
    return gen.eco_lite_deploy(segments=4, classes=7, batch=1)


def test_save_and_copy_from_round_trip(tmp_path):
    txt = small_net()
    ref = refnet.RefNet(txt).init_params(11)
    a = caffe.Net.from_string(txt, caffe.TEST)
    for name, arrs in ref.params_dict().items():
        for blob, arr in zip(a.params[name], arrs):
            blob.data[...] = arr
    path = str(tmp_path / "w.caffemodel")
    a.save(path)
    b = caffe.Net.from_string(txt, caffe.TEST)
    assert not np.array_equal(b.params["res3a_2n"][0].data, a.params["res3a_2n"][0].data)
    b.copy_from(path)
    for name, blobs in a.params.items():
        for x, y in zip(blobs, b.params[name]):
            assert np.array_equal(x.data, y.data), name
    # constructor form Net(prototxt, caffemodel, phase) and the RuntimeError on missing files (_caffe.cpp:57-64)
    proto = tmp_path / "deploy.prototxt"
    proto.write_text(txt)
    c = caffe.Net(str(proto), path, caffe.TEST)
    assert np.array_equal(c.params["fc8u"][1].data, a.params["fc8u"][1].data)
    with pytest.raises(RuntimeError):
        caffe.Net(str(proto), str(tmp_path / "missing.caffemodel"), caffe.TEST)
    with pytest.raises(RuntimeError):
        caffe.Net(str(tmp_path / "missing.prototxt"), caffe.TEST)


def test_copy_from_rejects_shape_mismatch_and_ignores_unknown_layers(tmp_path):
    a = caffe.Net.from_string(gen.eco_lite_deploy(segments=4, classes=7, batch=1), caffe.TEST)
    path = str(tmp_path / "w.caffemodel")
    a.save(path)
    # a net with another class count: fc layer shape differs -> error like caffe's "shape mismatch" CHECK
    b = caffe.Net.from_string(gen.eco_lite_deploy(segments=4, classes=9, batch=1), caffe.TEST)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        b.copy_from(path)
    # a net whose fc layer has a different NAME simply ignores the source layer (net.cpp:860-863)
    c = caffe.Net.from_string(gen.eco_lite_deploy(segments=4, classes=9, fc_name="fc8_other", batch=1), caffe.TEST)
    c.copy_from(path)
    assert np.array_equal(c.params["conv1_7x7_s2"][0].data, a.params["conv1_7x7_s2"][0].data)


PB2_SNIPPET = r'''
import sys, os
os.environ["PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION"] = "python"
sys.path.insert(0, %r)
import caffe_pb2, numpy as np
mode, path = sys.argv[1], sys.argv[2]
if mode == "read":
    net = caffe_pb2.NetParameter()
    net.ParseFromString(open(path, "rb").read())
    names = [l.name for l in net.layer]
    l = [l for l in net.layer if l.name == "res3a_2n"][0]
    w = np.array(l.blobs[0].data, np.float32)
    print("OK", len(names), l.type, list(l.blobs[0].shape.dim), "%%.6f" %% float(w.sum()), len(l.blobs))
else:
    net = caffe_pb2.NetParameter()
    net.name = "from_reference_schema"
    lay = net.layer.add(); lay.name = "conv1_7x7_s2"; lay.type = "Convolution"
    b = lay.blobs.add(); b.shape.dim.extend([64, 3, 7, 7]); b.data.extend(np.arange(64*3*7*7, dtype=np.float32) * 1e-4)
    b = lay.blobs.add(); b.shape.dim.extend([64]); b.data.extend(np.ones(64, np.float32) * 0.5)
    lay = net.layer.add(); lay.name = "conv1_7x7_s2_bn"; lay.type = "BN"          # legacy 4-D dims
    dims = (1, 64, 1, 1) if mode == "write_bad_legacy" else (1, 1, 1, 64)            # legacy blobs index from the END (blob.cpp:416-428)
    for v in (1.0, 2.0, 3.0, 4.0):
        b = lay.blobs.add(); b.num, b.channels, b.height, b.width = dims; b.data.extend(np.full(64, v, np.float32))
    lay = net.layer.add(); lay.name = "not_in_target_net"; lay.type = "Convolution"
    b = lay.blobs.add(); b.shape.dim.extend([2]); b.data.extend([1.0, 2.0])
    open(path, "wb").write(net.SerializeToString())
    print("OK")
'''


@pytest.mark.skipif(not os.path.isdir(REF_PROTO), reason="reference schema not mounted (GPU box)")
def test_wire_format_against_reference_schema(tmp_path):
    script = tmp_path / "pb2_tool.py"
    script.write_text(PB2_SNIPPET % REF_PROTO)
    txt = small_net()
    ref = refnet.RefNet(txt).init_params(5)
    a = caffe.Net.from_string(txt, caffe.TEST)
    for name, arrs in ref.params_dict().items():
        for blob, arr in zip(a.params[name], arrs):
            blob.data[...] = arr
    ours = str(tmp_path / "ours.caffemodel")
    a.save(ours)
    # (1) the reference's protobuf schema parses what we wrote
    out = subprocess.run([sys.executable, str(script), "read", ours], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1500:]
    tok = out.stdout.split()
    assert tok[0] == "OK" and int(tok[1]) == 109 and tok[2] == "Convolution"
    w = ref.params_dict()["res3a_2n"][0]
    assert out.stdout.strip().endswith("2") and ("[128, 96, 3, 3, 3]" in out.stdout)
    assert abs(float(tok[-2]) - float(w.sum())) < 1e-2
    # (2) we parse what the reference's schema writes (new-style shape, legacy 4-D dims, unknown layer)
    theirs = str(tmp_path / "theirs.caffemodel")
    out = subprocess.run([sys.executable, str(script), "write", theirs], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1500:]
    b = caffe.Net.from_string(txt, caffe.TEST)
    b.copy_from(theirs)
    assert np.allclose(b.params["conv1_7x7_s2"][0].data.ravel(), np.arange(64 * 3 * 7 * 7, dtype=np.float32) * 1e-4)
    assert np.all(b.params["conv1_7x7_s2"][1].data == 0.5)
    for k, v in enumerate((1.0, 2.0, 3.0, 4.0)):
        assert np.all(b.params["conv1_7x7_s2_bn"][k].data == v)
    # legacy dims are matched from the end, exactly as Blob::ShapeEquals does: (1,64,1,1) != a [1,64] blob
    bad = str(tmp_path / "bad.caffemodel")
    out = subprocess.run([sys.executable, str(script), "write_bad_legacy", bad], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-1500:]
    with pytest.raises(RuntimeError, match="shape mismatch"):
        caffe.Net.from_string(txt, caffe.TEST).copy_from(bad)
