"""Generates tests/golden/scan_statements.npz: point_cloud.cpp's scan-side functions executed STATEMENT BY
STATEMENT in Python with OpenCV doing what OpenCV does in the reference (cv::Mat products = cv2.gemm on
CV_64F, Mat::convertTo(CV_8U) = saturate_cast) and libm doing the rest (math.atan2 / sqrt / floor / tan are
the C library's).  The reference itself cannot be built here (ROS + OpenCV 2.4 C++), so this is the closest
executable form of its code: the same statements in the same order on a 64x48 crop.

  cacheDisparityValues()                       point_cloud.cpp:104-147
  publishObstacleScan(Mat& dmap, seq)          point_cloud.cpp:213-296   (default path)
  publishPointCloud() point loop               point_cloud.cpp:321-349   (-g path: points with d >= 2)
  publishObstacleScan(vector<Point3d>, seq)    point_cloud.cpp:149-211   (-g path: scan from the points)

Where the reference has undefined behaviour the script follows the definitions of DESIGN.md section 4: a bin
index outside [0, 89] (|theta| > 45 deg, point_cloud.cpp:264-267 writes outside scan[90]) is skipped, and so is
a NaN reprojection (d = 0 makes pos.w = 0).  Run in the build container only:
    python tests/golden/make_scan_statement_golden.py
"""
import json
import math
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
fx = json.load(open(os.path.join(HERE, "q_fixtures.json")))
Q = np.array(fx["Q"]["640x480"], np.float64)
XR = np.array(fx["calib"]["XR"], np.float64).reshape(3, 3)
XT = np.array(fx["calib"]["XT"], np.float64).reshape(3, 1)

INF = int(1e9)                                   # const int INF = 1e9;                       :55
GP_HEIGHT_THRESH = 0.05                          # :66
GP_ANGLE_THRESH = 4. * 3.1415 / 180.             # :67
GP_DIST_THRESH = 1.0                             # :68


def cache_disparity_values(crop_im_width, crop_im_height, crop_offset_x, crop_offset_y):
    valid_disp = np.zeros((crop_im_height, crop_im_width, 2), np.uint8)
    valid_disp[..., 0] = 255; valid_disp[..., 1] = 3          # Mat(h, w, CV_8UC2, Scalar(255,3))     :106
    V = np.zeros((4, 1), np.float64)
    for i in range(crop_im_width):                                                                  # :107
        for j in range(crop_im_height):                                                             # :108
            d = 3
            while d <= 255:                                                                         # :110
                V[0, 0] = float(i + crop_offset_x); V[1, 0] = float(j + crop_offset_y)
                V[2, 0] = float(d); V[3, 0] = 1.
                pos = cv2.gemm(Q, V, 1.0, None, 0.0)                                                # :115
                X = pos[0, 0] / pos[3, 0]; Y = pos[1, 0] / pos[3, 0]; Z = pos[2, 0] / pos[3, 0]
                point3d_cam = np.array([[X], [Y], [Z]], np.float64)
                point3d_robot = cv2.gemm(XR, point3d_cam, 1.0, XT, 1.0)                             # :123
                X = point3d_robot[0, 0]; Y = point3d_robot[1, 0]; Z = point3d_robot[2, 0]
                if Z < 0.:                                                                          # :128
                    d += 1
                    continue
                if X < GP_DIST_THRESH:                                                              # :133
                    if Z < GP_HEIGHT_THRESH:
                        d += 1
                        continue
                else:
                    if Z < GP_HEIGHT_THRESH + math.tan(GP_ANGLE_THRESH) * (X - GP_DIST_THRESH):     # :137
                        d += 1
                        continue
                break                                                                               # :140
            valid_disp[j, i, 0] = d & 0xFF       # uchar = int: 256 wraps to 0                       :142
            valid_disp[j, i, 1] = 255                                                               # :143
    return valid_disp


def reproject(i, j, d, ox, oy):
    V = np.array([[float(i + ox)], [float(j + oy)], [float(d)], [1.]], np.float64)
    pos = cv2.gemm(Q, V, 1.0, None, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        X = pos[0, 0] / pos[3, 0]; Y = pos[1, 0] / pos[3, 0]; Z = pos[2, 0] / pos[3, 0]
    p = cv2.gemm(XR, np.array([[X], [Y], [Z]], np.float64), 1.0, XT, 1.0)
    return p[0, 0], p[1, 0], p[2, 0]


def bin_point(scan, st, X, Y):
    fov = 90.; bin_size = 90
    if not (math.isfinite(X) and math.isfinite(Y)):
        return                                   # defined: NaN / inf reprojections are skipped
    theta_rad = math.atan2(Y, X)                                                                    # :255
    theta_deg = theta_rad * 180. / 3.1415                                                           # :256
    r = math.sqrt(Y * Y + X * X)                                                                    # :260
    k = math.floor(float(bin_size) * (fov / 2. - theta_deg) / fov)                                  # :264
    if k < 0 or k >= bin_size:
        return                                   # defined: the reference indexes outside scan[90] here
    st["min_angle"] = min(st["min_angle"], theta_rad); st["max_angle"] = max(st["max_angle"], theta_rad)
    st["range_max"] = max(st["range_max"], r); st["range_min"] = min(st["range_min"], r)
    st["n"] += 1
    if r < scan[k]:                                                                                 # :265
        scan[k] = r


def new_state():
    return {"min_angle": 400., "max_angle": -400., "range_min": float(INF), "range_max": -500., "n": 0}


def obstacle_scan_from_dmap(dmap, valid_disp, ox, oy):
    H, W = dmap.shape
    scan = [float(INF)] * 90                                                                        # :227-229
    st = new_state()
    for i in range(W):                                                                              # :230
        for j in range(H):                                                                          # :231
            d = int(dmap[j, i])
            if d < valid_disp[j, i, 0] or d > valid_disp[j, i, 1]:                                  # :234
                continue
            X, Y, Z = reproject(i, j, d, ox, oy)
            bin_point(scan, st, X, Y)
    return np.array(scan), st


def points_from_dmap(dmap, ox, oy):
    H, W = dmap.shape
    pts = []
    for i in range(W):                                                                              # :321
        for j in range(H):                                                                          # :322
            d = int(dmap[j, i])
            if d < 2:                                                                               # :326
                continue
            pts.append(reproject(i, j, d, ox, oy))
    return np.array(pts, np.float64).reshape(-1, 3)


def obstacle_scan_from_points(points):
    scan = [float(INF)] * 90
    st = new_state()
    for X, Y, Z in points:                                                                          # :164
        if Z < 0.:                                                                                  # :166
            continue
        if X < GP_DIST_THRESH:
            if Z < GP_HEIGHT_THRESH:
                continue
        else:
            if Z < GP_HEIGHT_THRESH + math.tan(GP_ANGLE_THRESH) * (X - GP_DIST_THRESH):
                continue
        bin_point(scan, st, X, Y)
    return np.array(scan), st


def compact(scan):
    return np.array([np.float32(scan[i]) for i in range(89, -1, -1) if scan[i] < INF - 1], np.float32)   # :278-282


out = {}
rng = np.random.default_rng(23)
cases = [("center", 64, 48, 300, 216, 18.0, "640x480"), ("left_edge", 64, 48, 0, 200, 40.0, "640x480"),
         ("low_right", 64, 48, 560, 400, 130.0, "640x480"), ("top", 64, 48, 100, 0, 2.0, "640x480"),
         ("bottom", 64, 48, 300, 432, 150.0, "640x480"),
         # the bench geometry (K x 3 at 1920x1200): near the lower border no disparity <= 255 clears the ground
         # gate, `d` leaves the loop as 256 and wraps to 0 in the Vec2b (H8) -- every disparity passes there
         ("kx3_bottom", 64, 48, 900, 1150, 100.0, "1920x1200_Kx3")]
for name, W, H, ox, oy, d0, qk in cases:
    Q = np.array(fx["Q"][qk], np.float64)
    gate = cache_disparity_values(W, H, ox, oy)
    # a disparity map as generateDisparityMap leaves it: float -> u8 by convertTo (saturate_cast, :422)
    yy, xx = np.mgrid[0:H, 0:W]
    D = (d0 + 0.9 * yy + 0.05 * xx + rng.normal(0, 0.4, (H, W))).astype(np.float32)
    D[rng.random((H, W)) < 0.15] = -10.0
    D[rng.random((H, W)) < 0.05] = rng.uniform(0, 255, 1).astype(np.float32)[0]
    D[H // 3:H // 2, W // 4:W // 2] = 121.5
    u8 = cv2.add(D, np.zeros_like(D), dtype=cv2.CV_8U)
    scan, st = obstacle_scan_from_dmap(u8, gate, ox, oy)
    pts = points_from_dmap(u8, ox, oy)
    scan_p, st_p = obstacle_scan_from_points(pts)
    out.update({name + "_" + k: v for k, v in dict(
        dims=np.array([W, H, ox, oy]), Q=Q, gate=gate, D=D, u8=u8, scan=scan, meta=np.array(
            [st["min_angle"], st["max_angle"], st["range_min"], st["range_max"], st["n"]]),
        compact=compact(scan), pts=pts, scan_p=scan_p, meta_p=np.array(
            [st_p["min_angle"], st_p["max_angle"], st_p["range_min"], st_p["range_max"], st_p["n"]])).items()})
    print(name, "gate min/max", gate[..., 0].min(), gate[..., 0].max(), "wrapped to 0:", int((gate[..., 0] == 0).sum()),
          "points in scan", st["n"], "finite bins", int((scan < INF - 1).sum()), "cloud points", len(pts),
          "scan-from-points bins", int((scan_p < INF - 1).sum()))
np.savez_compressed(os.path.join(HERE, "scan_statements.npz"), XR=XR, XT=XT, names=np.array([c[0] for c in cases]),
                    cv_version=np.array(cv2.__version__), **out)
print("scan_statements.npz", os.path.getsize(os.path.join(HERE, "scan_statements.npz")) // 1024, "KiB")
