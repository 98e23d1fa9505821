"""Generates tests/golden/stereo_rectify_cv2.npz: cv2.stereoRectify / cv2.initUndistortRectifyMap (OpenCV 4.13)
outputs for the shipped calibration in the reference's call pattern (point_cloud.cpp:543-554) and for
perturbed calibrations (vertical stereo, no ZERO_DISPARITY flag, other alpha), against which the OpenCV-free
host code in csrc/calib.cu is checked.  Maps are stored on a sub-sampled grid.  Build container only:
    python tests/golden/make_rectify_golden.py
"""
import json
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
fx = json.load(open(os.path.join(HERE, "q_fixtures.json")))["calib"]
K1 = np.array(fx["K1"]); K2 = np.array(fx["K2"]); D1 = np.array(fx["D1"]).reshape(-1); D2 = np.array(fx["D2"]).reshape(-1)
R = np.array(fx["R"]); T = np.array(fx["T"])
rng = np.random.default_rng(9)
cases = []
for newsize, scale in (((640, 480), 1.0), ((320, 180), 1.0), ((1920, 1200), 3.0), ((640, 360), 1.0)):
    k1 = K1.copy(); k2 = K2.copy(); k1[:2] *= scale; k2[:2] *= scale
    cases.append(dict(K1=k1, K2=k2, D1=D1, D2=D2, R=R, T=T, size=(int(640 * scale), int(360 * scale)), newsize=newsize,
                      zero=1, alpha=0.0))
for i in range(8):
    k1 = K1 * (1 + rng.uniform(-0.05, 0.05)); k1[2, 2] = 1
    k2 = K2 * (1 + rng.uniform(-0.05, 0.05)); k2[2, 2] = 1
    d1 = D1 * rng.uniform(0.5, 1.5, 5); d2 = D2 * rng.uniform(0.5, 1.5, 5)
    rv = cv2.Rodrigues(R)[0].reshape(3) + rng.uniform(-0.03, 0.03, 3)
    r = cv2.Rodrigues(rv)[0]
    t = T * rng.uniform(0.7, 1.3) + rng.uniform(-0.005, 0.005, 3)
    if i % 4 == 3:
        t = np.array([t[1], t[0], t[2]])                  # vertical stereo
    cases.append(dict(K1=k1, K2=k2, D1=d1, D2=d2, R=r, T=t, size=(640, 360), newsize=[(640, 360), (800, 600), (0, 0)][i % 3],
                      zero=int(i % 2 == 0), alpha=[0.0, -1.0, 0.5, 1.0][i % 4]))
out = {"n": np.array(len(cases))}
for i, c in enumerate(cases):
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(c["K1"], c["D1"], c["K2"], c["D2"], c["size"], c["R"], c["T"],
                                                flags=cv2.CALIB_ZERO_DISPARITY if c["zero"] else 0, alpha=c["alpha"],
                                                newImageSize=tuple(c["newsize"]))
    w, h = c["newsize"] if c["newsize"][0] else c["size"]
    mx, my = cv2.initUndistortRectifyMap(c["K1"], c["D1"], R1, P1, (w, h), cv2.CV_32F)
    mx2, my2 = cv2.initUndistortRectifyMap(c["K2"], c["D2"], R2, P2, (w, h), cv2.CV_32F)
    for k, v in dict(K1=c["K1"], K2=c["K2"], D1=c["D1"], D2=c["D2"], R=c["R"], T=c["T"],
                     cfg=np.array([c["size"][0], c["size"][1], c["newsize"][0], c["newsize"][1], c["zero"]], np.int32),
                     alpha=np.array(c["alpha"]), R1=R1, R2=R2, P1=P1, P2=P2, Q=Q,
                     mx=mx[::23, ::29], my=my[::23, ::29], mx2=mx2[::23, ::29], my2=my2[::23, ::29]).items():
        out["c%d_%s" % (i, k)] = np.asarray(v)
    print(i, c["size"], c["newsize"], "zero", c["zero"], "alpha", c["alpha"], "f %.3f" % P1[0, 0])
np.savez_compressed(os.path.join(HERE, "stereo_rectify_cv2.npz"), cv_version=np.array(cv2.__version__), **out)
print(os.path.getsize(os.path.join(HERE, "stereo_rectify_cv2.npz")) // 1024, "KiB")
