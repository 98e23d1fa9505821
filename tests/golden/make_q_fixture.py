"""Generates tests/golden/q_fixtures.json: the reprojection matrix Q that the reference
obtains from cv::stereoRectify (point_cloud.cpp:543-544) for the shipped calibration
file, computed here with OpenCV's Python binding (cv2 4.13; the reference's OpenCV 2.4 is
not available -- parity at this boundary is unpinned).

Run in the build container only (needs cv2 and /root/reference):
    python tests/golden/make_q_fixture.py
"""
import json
import os
import cv2
import numpy as np

REF = "/root/reference/calibration/amrl_jackal_webcam_stereo.yml"
fs = cv2.FileStorage(REF, cv2.FILE_STORAGE_READ)
K1 = fs.getNode("K1").mat(); K2 = fs.getNode("K2").mat()
D1 = fs.getNode("D1").mat(); D2 = fs.getNode("D2").mat()
R = fs.getNode("R").mat()
T = np.array([fs.getNode("T").at(i).real() for i in range(3)], np.float64)
XR = fs.getNode("XR").mat(); XT = fs.getNode("XT").mat()
out = {"calib": {"K1": K1.tolist(), "K2": K2.tolist(), "D1": D1.tolist(), "D2": D2.tolist(), "R": R.tolist(),
                 "T": T.tolist(), "XR": XR.tolist(), "XT": XT.tolist()}, "Q": {}}
# reference call: stereoRectify(K1,D1,K2,D2, calib_im_size=(640,360), R, T, ..., CALIB_ZERO_DISPARITY, alpha=0,
#                               newImageSize=rawimsize)
for name, (w, h), scale in (("640x480", (640, 480), 1.0), ("320x180", (320, 180), 1.0),
                            ("1920x1200_Kx3", (1920, 1200), 3.0)):
    k1 = K1.copy(); k2 = K2.copy()
    k1[:2] *= scale; k2[:2] *= scale
    calib_size = (int(640 * scale), int(360 * scale))
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(k1, D1, k2, D2, calib_size, R, T, flags=cv2.CALIB_ZERO_DISPARITY,
                                                alpha=0, newImageSize=(w, h))
    out["Q"][name] = Q.tolist()
    print(name, "cx %.4f cy %.4f f %.4f q32 %.6f q33 %.6f" % (-Q[0, 3], -Q[1, 3], Q[2, 3], Q[3, 2], Q[3, 3]))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "q_fixtures.json"), "w"), indent=1)


def _mat(name, m):
    m = np.asarray(m, np.float64)
    r, c = m.shape
    vals = ", ".join(repr(float(x)) for x in m.reshape(-1))
    return "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ %s ]\n" % (name, r, c, vals)


# Calibration fixture in the reference's YAML schema (values of the shipped C920 calibration,
# re-serialised; T stays a bare 3-sequence as in the original file).
with open(os.path.join(os.path.dirname(__file__), "calib_c920.yml"), "w") as f:
    f.write("%YAML:1.0\n")
    f.write(_mat("K1", K1)); f.write(_mat("K2", K2)); f.write(_mat("D1", D1)); f.write(_mat("D2", D2))
    f.write(_mat("R", R))
    f.write("T: [ %s ]\n" % ", ".join(repr(float(x)) for x in T))
    f.write(_mat("XR", XR)); f.write(_mat("XT", XT))
