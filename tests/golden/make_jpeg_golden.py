"""Generates tests/golden/jpeg_cv2.npz: JPEG bitstreams (cv2.imencode, grayscale and colour, two qualities)
of synthetic textured frames together with what the reference's decode call returns for them,
cv::imdecode(data, CV_LOAD_IMAGE_GRAYSCALE) (point_cloud.cpp:436) -- here cv2.imdecode(..., IMREAD_GRAYSCALE)
of OpenCV 4.13 (libjpeg-turbo).  Build container only:  python tests/golden/make_jpeg_golden.py
"""
import importlib
import os
import sys
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
synth = importlib.import_module("jackal-navigation_b200.synth")
out = {}
n = 0
for seed, color, q in ((1, False, 95), (2, False, 80), (3, True, 95), (4, True, 70)):
    I1, I2, _ = synth.textured_pair(320, 240, 64, seed)
    img = I1 if not color else np.stack([I1, I2, np.roll(I1, 7, axis=1)], -1)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
    assert ok
    dec = cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)
    out["jpg%d" % n] = np.frombuffer(buf.tobytes(), np.uint8)
    out["gray%d" % n] = dec
    print(n, "color" if color else "gray", "q", q, len(buf), "bytes")
    n += 1
np.savez_compressed(os.path.join(HERE, "jpeg_cv2.npz"), n=np.array(n), cv_version=np.array(cv2.__version__), **out)
print(os.path.getsize(os.path.join(HERE, "jpeg_cv2.npz")) // 1024, "KiB")
