"""The C++ facade (include/eco_caffe_facade.hpp) and the C ABI link and behave like the caffe_3d
surface they mirror -- CPU part: build a net, names, shapes, legacy accessors; no compute without a GPU."""
import os
import re
import subprocess
import sys
import tempfile

import pytest

import gen_eco_prototxt as gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "eco-efficient-video-understanding_b200", "lib")

SRC = r'''
#define ECO_FACADE_THROW
#include "eco_caffe_facade.hpp"
#include <cstdio>
int main(int argc, char** argv) {
  caffe::Net<float> net(argv[1], caffe::TEST);
  std::printf("name=%s layers=%zu blobs=%zu in=%d out=%d\n", net.name().c_str(), net.layer_names().size(),
              net.blob_names().size(), net.num_inputs(), net.num_outputs());
  auto fc8 = net.blob_by_name("fc8");
  std::printf("fc8 %d x %d\n", fc8->shape(0), fc8->shape(1));
  auto v = net.blob_by_name("res2b_bn");
  std::printf("res2b_bn axes=%d count=%d\n", v->num_axes(), v->count());
  try { v->num(); std::printf("legacy ok\n"); } catch (const std::exception& e) { std::printf("legacy: %s\n", e.what()); }
  std::printf("has res3a split: %d unknown: %d\n", (int)net.has_layer("res3a_res3a_2n_0_split"), (int)net.has_blob("nope"));
  auto conv1 = net.layer_by_name("conv1_7x7_s2");
  std::printf("conv1 type=%s blobs=%zu w=%dx%dx%dx%d params=%zu layers=%zu\n", conv1->type(), conv1->blobs().size(),
              conv1->blobs()[0]->num(), conv1->blobs()[0]->channels(), conv1->blobs()[0]->height(), conv1->blobs()[0]->width(),
              net.params().size(), net.layers().size());
  conv1->blobs()[1]->mutable_cpu_data()[3] = 0.25f;
  std::printf("bias[3]=%.2f asum>0:%d offset=%d\n", conv1->blobs()[1]->cpu_data()[3], (int)(conv1->blobs()[0]->asum_data() > 0),
              conv1->blobs()[0]->offset(1, 2, 3, 4));
  caffe::Net<float> twin(argv[1], caffe::TEST);
  twin.ShareTrainedLayersWith(&net);
  std::printf("shared bias[3]=%.2f\n", twin.layer_by_name("conv1_7x7_s2")->blobs()[1]->cpu_data()[3]);
  float* in = net.input_blobs()[0]->mutable_cpu_data();
  in[0] = 1.f;
  try { net.ForwardPrefilled(); std::printf("forward ok\n"); }
  catch (const std::exception& e) { std::printf("forward: %s\n", e.what()); }
  return 0;
}
'''


def test_c_abi_exports_every_declared_symbol():
    import ctypes
    lib = ctypes.CDLL(os.path.join(LIBDIR, "libeco_b200.so"))
    hdr = open(os.path.join(ROOT, "include", "eco_b200.h")).read()
    syms = sorted(set(re.findall(r"\b(eco_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), s


def test_facade_compiles_and_introspects(tmp_path):
    src = tmp_path / "facade_demo.cpp"
    src.write_text(SRC)
    proto = tmp_path / "deploy.prototxt"
    proto.write_text(gen.eco_lite_deploy(16, 101, batch=1))
    exe = tmp_path / "facade_demo"
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", LIBDIR, "-leco_b200", "-Wl,-rpath," + LIBDIR], env=env)
    out = subprocess.run([str(exe), str(proto)], capture_output=True, text=True, timeout=120).stdout
    assert "name=o3d layers=116 blobs=97 in=1 out=1" in out
    assert "fc8 1 x 101" in out
    assert "res2b_bn axes=5 count=%d" % (96 * 16 * 28 * 28) in out
    assert "legacy: " in out and "legacy accessors" in out     # blob.hpp:141 behaviour on 5-D blobs
    assert "has res3a split: 1 unknown: 0" in out
    assert "conv1 type=Convolution blobs=2 w=64x3x7x7 params=186 layers=116" in out, out
    assert "bias[3]=0.25 asum>0:1 offset=%d" % (((1 * 3 + 2) * 7 + 3) * 7 + 4) in out
    assert "shared bias[3]=0.25" in out
    assert ("forward ok" in out) or ("no CUDA device" in out)    # loud failure without a GPU, never a CPU fallback
