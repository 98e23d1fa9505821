"""Generates tests/golden/scan_cv2.npz: the pieces of point_cloud.cpp's per-pixel arithmetic that live in
OpenCV, evaluated by OpenCV itself (cv2 4.13) so that oracle/scan_port.c can be pinned against them:

  * `pos = Q * V` and `XR * point3d_cam + XT` (point_cloud.cpp:241, 250, 336, 346) are cv::Mat
    expressions = cv::gemm on CV_64F; here cv2.gemm(Q, V, 1, None, 0) and cv2.gemm(XR, p, 1, XT, 1);
  * `Mat::convertTo(CV_8U)` (point_cloud.cpp:422) = saturate_cast<uchar>(float) per element; here
    cv2.add(D, 0, dtype=CV_8U), which applies the same saturate_cast to the float sum D + 0.

atan2 / sqrt / floor are libm and not part of the fixture.  Run in the build container only:
    python tests/golden/make_scan_golden.py
"""
import json
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
fx = json.load(open(os.path.join(HERE, "q_fixtures.json")))
Q = np.array(fx["Q"]["640x480"], np.float64)
XR = np.array(fx["calib"]["XR"], np.float64).reshape(3, 3)
XT = np.array(fx["calib"]["XT"], np.float64).reshape(3, 1)

rng = np.random.default_rng(17)
W, H, ox, oy = 48, 40, 7, 3
dmap = rng.integers(0, 256, (H, W), dtype=np.uint8)
dmap[rng.random((H, W)) < 0.2] = rng.integers(0, 3)          # some below the d >= 2 gate of the -g path
pts = []
for i in range(W):                                            # columns outer, like point_cloud.cpp:321-322
    for j in range(H):
        d = int(dmap[j, i])
        if d < 2:
            continue
        V = np.array([[i + ox], [j + oy], [d], [1.0]], np.float64)
        pos = cv2.gemm(Q, V, 1.0, None, 0.0)
        p = np.array([[pos[0, 0] / pos[3, 0]], [pos[1, 0] / pos[3, 0]], [pos[2, 0] / pos[3, 0]]], np.float64)
        pr = cv2.gemm(XR, p, 1.0, XT, 1.0)
        pts.append(pr.reshape(3))
pts = np.array(pts, np.float64)

# a general (dense) Q as well: the association order of the 4-term dot products matters there
Qg = Q + rng.uniform(-1e-3, 1e-3, (4, 4))
ptsg = []
for i in range(0, W, 3):
    for j in range(0, H, 3):
        d = int(dmap[j, i])
        if d < 2:
            continue
        V = np.array([[i + ox], [j + oy], [d], [1.0]], np.float64)
        pos = cv2.gemm(Qg, V, 1.0, None, 0.0)
        p = np.array([[pos[0, 0] / pos[3, 0]], [pos[1, 0] / pos[3, 0]], [pos[2, 0] / pos[3, 0]]], np.float64)
        ptsg.append(cv2.gemm(XR, p, 1.0, XT, 1.0).reshape(3))
ptsg = np.array(ptsg, np.float64)

# float -> u8: ties, negatives, the -10 / -1 markers, values beyond 255, halves from the mean filter
D = np.concatenate([np.arange(-12, 300, 0.25, dtype=np.float32), rng.uniform(-20, 300, 4000).astype(np.float32),
                    np.float32([-10, -1, 0.5, 1.5, 2.5, 254.5, 255.5, 1e9, -1e9])]).reshape(1, -1)
u8 = cv2.add(D, np.zeros_like(D), dtype=cv2.CV_8U)

np.savez_compressed(os.path.join(HERE, "scan_cv2.npz"), Q=Q, Qg=Qg, XR=XR, XT=XT, dmap=dmap, ox=ox, oy=oy, pts=pts,
                    ptsg=ptsg, D=D, u8=u8, cv_version=np.array(cv2.__version__))
print("scan_cv2.npz", os.path.getsize(os.path.join(HERE, "scan_cv2.npz")) // 1024, "KiB;", len(pts), "+", len(ptsg), "points")
