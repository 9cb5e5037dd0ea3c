#!/usr/bin/env python
"""Extract the known-answer vectors that the reference's own gtest suite holds for
the pooling layers on ECO's path and write them to reference_known_answers.json.

Runs only where /root/reference is mounted (this container); the JSON it writes is
committed and is what tests/test_oracle_golden.py (CPU) and tests/test_gpu_ops.py
(GPU) check against.  Nothing is copied from the reference except the numeric
literals of its test fixtures, located by line range:

  caffe_3d/src/caffe/test/test_pooling_layer.cpp
     :46-116    MAX 2x2 on 3x5           :118-241   MAX 3x2 (kernel_h 3, kernel_w 2) on 6x6
     :243-371   MAX 2x3 on 6x6           :475-518   MAX 3/s2/p2 on 3x3 (padded)
     :540-570   AVE 3/s1/p1 on const 2   :1221-1253 3-D fixture input 4x3x6
     :1255-1288 3-D MAX 2x2x2            :1290-1325 3-D MAX 2x2x3
     :1327-1359 3-D MAX 2x3x2            :1361-1394 3-D MAX 3x2x2
     :1469-1525 3-D MAX 3/s2/p2 padded   :1559-1613 3-D AVE 3/s1/p1 (27 values, tol 1e-4)
"""
import json
import os
import re

REF = "/root/reference/caffe_3d/src/caffe/test/test_pooling_layer.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")


def lines(a, b):
    with open(REF) as f:
        src = f.read().split("\n")
    return "\n".join(src[a - 1:b])


def array_literal(text, name):
    m = re.search(r"%s\[\]\s*=\s*\{(.*?)\};" % name, text, re.S)
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return [float(t) for t in re.findall(r"[-+]?\d+\.?\d*", body)]


def indexed_assign(text):
    vals = {}
    for m in re.finditer(r"mutable_cpu_data\(\)\[(?:i\s*\+\s*)?(\d+)\]\s*=\s*([-\d.]+);", text):
        vals[int(m.group(1))] = float(m.group(2))
    return [vals[i] for i in range(len(vals))]


def indexed_expect(text, blob="blob_top_"):
    vals = {}
    pat = r"EXPECT_(?:EQ|NEAR)\((?:this->)?%s->cpu_data\(\)\[(?:i\s*\+\s*)?(\d+)\],\s*([^,)]+?)\s*[,)]" % blob
    for m in re.finditer(pat, text):
        vals[int(m.group(1))] = float(eval(m.group(2), {"__builtins__": {}}))
    return [vals[i] for i in range(len(vals))]


def main():
    cases = []
    # --- 2-D MAX, exact integers ------------------------------------------------
    t = lines(46, 116)
    cases.append(dict(name="max_2x2_on_3x5", cite="test_pooling_layer.cpp:46-116", method="MAX",
                      in_shape=[3, 5], kernel=[2, 2], stride=[1, 1], pad=[0, 0],
                      input=indexed_assign(t), out_shape=[2, 4], output=indexed_expect(t), tol=0))
    t = lines(118, 241)
    cases.append(dict(name="max_3x2_on_6x6", cite="test_pooling_layer.cpp:118-241", method="MAX",
                      in_shape=[6, 6], kernel=[3, 2], stride=[1, 1], pad=[0, 0],
                      input=indexed_assign(t), out_shape=[4, 5], output=indexed_expect(t), tol=0))
    t = lines(243, 371)
    cases.append(dict(name="max_2x3_on_6x6", cite="test_pooling_layer.cpp:243-371", method="MAX",
                      in_shape=[6, 6], kernel=[2, 3], stride=[1, 1], pad=[0, 0],
                      input=indexed_assign(t), out_shape=[5, 4], output=indexed_expect(t), tol=0))
    t = lines(475, 518)
    cases.append(dict(name="max_3_s2_p2_on_3x3", cite="test_pooling_layer.cpp:475-518", method="MAX",
                      in_shape=[3, 3], kernel=[3, 3], stride=[2, 2], pad=[2, 2],
                      input=indexed_assign(t), out_shape=[3, 3], output=indexed_expect(t), tol=1e-8))
    t = lines(540, 570)
    cases.append(dict(name="ave_3_s1_p1_const2", cite="test_pooling_layer.cpp:540-570", method="AVE",
                      in_shape=[3, 3], kernel=[3, 3], stride=[1, 1], pad=[1, 1],
                      input=[2.0] * 9, out_shape=[3, 3], output=indexed_expect(t), tol=1e-5))
    # --- 3-D (the reference can only run these through cuDNN) --------------------
    fix = array_literal(lines(1221, 1253), "input")
    for a, b, k, osh, nm in ((1255, 1288, [2, 2, 2], [3, 2, 5], "cube"),
                             (1290, 1325, [2, 2, 3], [3, 2, 4], "cuboid_x"),
                             (1327, 1359, [2, 3, 2], [3, 1, 5], "cuboid_y"),
                             (1361, 1394, [3, 2, 2], [2, 2, 5], "cuboid_z")):
        t = lines(a, b)
        cases.append(dict(name="max3d_" + nm, cite="test_pooling_layer.cpp:%d-%d" % (a, b), method="MAX",
                          in_shape=[4, 3, 6], kernel=k, stride=[1, 1, 1], pad=[0, 0, 0],
                          input=fix, out_shape=osh, output=array_literal(t, "output"), tol=0))
    t = lines(1469, 1525)
    cases.append(dict(name="max3d_3_s2_p2", cite="test_pooling_layer.cpp:1469-1525", method="MAX",
                      in_shape=[3, 3, 3], kernel=[3, 3, 3], stride=[2, 2, 2], pad=[2, 2, 2],
                      input=array_literal(t, "input"), out_shape=[3, 3, 3],
                      output=array_literal(t, "output"), tol=1e-8))
    t = lines(1559, 1613)
    cases.append(dict(name="ave3d_3_s1_p1", cite="test_pooling_layer.cpp:1559-1613", method="AVE",
                      in_shape=[3, 3, 3], kernel=[3, 3, 3], stride=[1, 1, 1], pad=[1, 1, 1],
                      input=array_literal(t, "input"), out_shape=[3, 3, 3],
                      output=array_literal(t, "output"), tol=1e-4))
    for c in cases:
        n_in = 1
        for d in c["in_shape"]:
            n_in *= d
        n_out = 1
        for d in c["out_shape"]:
            n_out *= d
        assert len(c["input"]) == n_in, (c["name"], len(c["input"]), n_in)
        assert len(c["output"]) == n_out, (c["name"], len(c["output"]), n_out)
    with open(OUT, "w") as f:
        json.dump({"source": "caffe_3d/src/caffe/test/test_pooling_layer.cpp", "cases": cases}, f, indent=1)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
