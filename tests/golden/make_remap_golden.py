"""Generates tests/golden/remap_cv2.npz: outputs of OpenCV's own cv2.remap (the library the reference
calls at point_cloud.cpp:440,481) for the rectification maps of the shipped calibration file and for
adversarial random maps.  Run in the build container only (needs cv2):

    python tests/golden/make_remap_golden.py
"""
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}

# (a) the reference's own init sequence (point_cloud.cpp:530-554) on the shipped calibration, at a
#     reduced raw image size so that the fixture stays small
fs = cv2.FileStorage(os.path.join(HERE, "calib_c920.yml"), cv2.FILE_STORAGE_READ)
K1, K2 = fs.getNode("K1").mat(), fs.getNode("K2").mat()
D1, D2 = fs.getNode("D1").mat(), fs.getNode("D2").mat()
R = fs.getNode("R").mat()
T = np.array([fs.getNode("T").at(i).real() for i in range(3)], np.float64).reshape(3, 1)
calib_size, raw = (640, 360), (160, 90)
R1, R2, P1, P2, Q, roi1, roi2 = cv2.stereoRectify(K1, D1, K2, D2, calib_size, R, T, flags=cv2.CALIB_ZERO_DISPARITY,
                                                  alpha=0, newImageSize=raw)
rng = np.random.default_rng(7)
for side, (K, D, Rr, P) in enumerate(((K1, D1, R1, P1), (K2, D2, R2, P2))):
    mx, my = cv2.initUndistortRectifyMap(K, D, Rr, P, raw, cv2.CV_32F)
    # camera frame at calibration size; regenerated from its seed by the tests (not stored)
    src = np.random.default_rng(100 + side).integers(0, 256, (360, 640), dtype=np.uint8)
    out["calib%d_mapx" % side] = mx
    out["calib%d_mapy" % side] = my
    out["calib%d_src_seed" % side] = np.array([100 + side, 360, 640])
    out["calib%d_dst" % side] = cv2.remap(src, mx, my, cv2.INTER_LINEAR)

# (b) adversarial maps: exact integers, halves (round-half-even of x*32), negatives, far outside,
#     the last row/column, and coordinates straddling every border
H, W, SH, SW = 96, 128, 75, 101
src = rng.integers(0, 256, (SH, SW), dtype=np.uint8)
mx = rng.uniform(-3, SW + 3, (H, W)).astype(np.float32)
my = rng.uniform(-3, SH + 3, (H, W)).astype(np.float32)
mx[0:8] = np.round(mx[0:8])                                      # integer x
my[4:12] = np.round(my[4:12])                                    # integer y (rows 4..7: both)
mx[12:20] = (np.round(mx[12:20] * 64) / 64).astype(np.float32)   # multiples of 1/64: ties of cvRound(x*32)
my[16:24] = (np.round(my[16:24] * 64) / 64).astype(np.float32)
mx[24, :8] = [-1, -0.5, -1.0 / 64, 0, SW - 1, SW - 1 + 1.0 / 64, SW - 0.5, SW]
my[25, :8] = [-1, -0.5, -1.0 / 64, 0, SH - 1, SH - 1 + 1.0 / 64, SH - 0.5, SH]
mx[26, :4] = [-1e4, 1e4, -40000, 40000]
my[27, :4] = [-1e4, 1e4, -40000, 40000]
out["rand_mapx"], out["rand_mapy"], out["rand_src"] = mx, my, src
out["rand_dst"] = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
out["cv_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(HERE, "remap_cv2.npz"), **out)
print("remap_cv2.npz", os.path.getsize(os.path.join(HERE, "remap_cv2.npz")) // 1024, "KiB; valid ROI", roi1, roi2)
