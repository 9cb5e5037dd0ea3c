#!/usr/bin/env python
"""Golden vectors for the resize step of DataTransformer::Transform (data_transformer.cpp:269: cv::resize(cropM, buffer,
cv::Size(crop_size, crop_size)), default INTER_LINEAR, 8-bit planes), generated with the real OpenCV of this image
(the reference does not pin an OpenCV version).  Writes tests/golden/resize_cv2.npz."""
import os
import cv2
import numpy as np

rng = np.random.default_rng(2024)
cases = {}
for i, (sh, sw, d) in enumerate([(256, 256, 224), (224, 196, 224), (196, 224, 224), (168, 168, 224), (224, 168, 224),
                                 (31, 47, 24), (24, 24, 24), (17, 40, 32)]):
    src = rng.integers(0, 256, size=(sh, sw), dtype=np.uint8)
    cases["src%d" % i] = src
    cases["dst%d" % i] = cv2.resize(src, (d, d))
cases["opencv_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_cv2.npz"), **cases)
print("wrote", len(cases) // 2, "cases with OpenCV", cv2.__version__)
