#!/usr/bin/env python
"""What examples/camera_to_command.c must print, computed WITHOUT the library's device code: the program's frames
(same LCG, same painting order) through the compiled reference ELAS (oracle/_ref), the scan restatement
(oracle/scan_port.c) and the vote (host code).  Test infrastructure; takes about a minute (pure-Python frame loop)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import scan_lib  # noqa: E402

W, H = 640, 360


def make_pair(box_d, seed):
    n = 2 * W * H
    out = np.empty(n, np.uint32)
    x = seed
    for i in range(n):
        x = (1664525 * x + 1013904223) & 0xFFFFFFFF
        out[i] = x
    v = ((out >> 8) & 255).astype(np.uint8)
    L, R = v[0::2].reshape(H, W).copy(), v[1::2].reshape(H, W).copy()
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    in_box = (uu > W // 8) & (uu < W // 2) & (vv > H // 4) & (vv < 3 * H // 4)
    d = np.where(in_box, box_d, 6 + 40 * vv // H)
    for p in (False, True):                       # far surface first, the box over it; within a pass u ascending
        for r in range(H):
            sel = (in_box[r] == p) & (uu[r] - d[r] >= 0)
            us = uu[r][sel]
            R[r, us - d[r][sel]] = L[r, us]       # targets within a row and pass are distinct: d is constant there
    return L, R


if __name__ == "__main__":
    jn = importlib.import_module("jackal-navigation_b200")
    o = ol.load("ref") or ol.load("port")
    sp = scan_lib.ScanPort()
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.stereo_rectify(640, 360, W, H)
    A = cal.arrays()
    gate = sp.gate(A["Q"], A["XR"], A["XT"], W, H)
    nav = jn.Navigate()
    for f in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        L, R = make_pair(30 + 25 * f, 1000 + f)
        D1, _ = o.process(ol.robotics(255, postprocess_only_left=1), L, R)
        r, m = sp.scan(A["Q"], A["XR"], A["XT"], gate, sp.convert_u8(D1))
        nav.set_scan_bins(r, jn.ScanMeta(m.angle_min, m.angle_max, m.range_min, m.range_max, m.n_finite, m.n_points))
        v = nav.command(jn.NAV_OBSTACLE_AVOID, 0.0, 1.0)
        print("frame %d: %d scan bins, nearest %.2f m -> cmd_vel linear %.3f m/s, angular %+.3f rad/s (%s)" % (
            f, m.n_finite, m.range_min, v[0], v[1],
            "turning left" if v[1] > 0 else "turning right" if v[1] < 0 else "driving" if v[0] > 0 else "stopped"))
