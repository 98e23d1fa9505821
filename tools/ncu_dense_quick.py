#!/usr/bin/env python
"""Smallest possible host for one `ncu --set full` capture of dense_kernel: no torch, no CPU checker.

    python tools/ncu_dense_quick.py --make          # here: writes tools/sweep/ncu_frames.npz (4 pairs, seeds 1000-1003)
    JN_ELAS_SPLIT=1 ncu --set full --clock-control none -k regex:dense_kernel -c 1 -o gpurun_out/rep \
        python tools/ncu_dense_quick.py             # on the GPU box: one 4-frame launch of every ELAS kernel

Four 1920x1200 pairs go through jn_stereo_scan_submit (host buffers in, scans + D1 out); the D1 digests are printed
so that the run can be checked against the compiled reference afterwards (tools/ncu_dense_quick.py --check).
"""
import hashlib
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
FRAMES = os.path.join(ROOT, "tools", "sweep", "ncu_frames.npz")
W, H, DM, SEEDS = 1920, 1200, 255, [1000, 1001, 1002, 1003]


def digest(a):
    return hashlib.sha1(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


def make():
    synth = importlib.import_module("jackal-navigation_b200.synth")
    L = np.empty((len(SEEDS), H, W), np.uint8)
    R = np.empty_like(L)
    for i, s in enumerate(SEEDS):
        L[i], R[i], _ = synth.SCENES["random_dot"](W, H, DM, s)
    os.makedirs(os.path.dirname(FRAMES), exist_ok=True)
    np.savez(FRAMES, L=L, R=R)
    print("wrote", FRAMES)


def check():
    import oracle_lib as ol
    z = np.load(FRAMES)
    o = ol.load("ref")
    for i in range(len(SEEDS)):
        D1, _ = o.process(ol.robotics(DM), z["L"][i], z["R"][i])
        print("ref D1[%d] %s" % (i, digest(D1)))


def run():
    t0 = time.time()
    import json
    jn = importlib.import_module("jackal-navigation_b200")
    z = np.load(FRAMES)
    L, R = np.ascontiguousarray(z["L"]), np.ascontiguousarray(z["R"])
    n = L.shape[0]
    Q = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "q_fixtures.json")))["Q"]["1920x1200_Kx3"])
    cal = jn.Calibration(os.path.join(ROOT, "tests", "golden", "calib_c920.yml"))
    cal.set_q_matrix(Q)
    sc = jn.ObstacleScan(cal, W, H)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=DM))
    D1 = np.zeros((n, H, W), np.float32)
    st = np.full(n, -9, np.int32)
    ranges = np.zeros((n, 90), np.float64)
    meta = np.zeros((n, 5), np.float64)
    print("setup %.2fs" % (time.time() - t0), flush=True)
    for rep in range(2):
        e.stereo_scan_submit(sc, n, L.ctypes.data, R.ctypes.data, (W, H, W), ranges.ctypes.data, meta.ctypes.data,
                             st.ctypes.data, 0, D1.ctypes.data)
        e.stereo_scan_wait()
        print("pass %d done %.2fs status %s bins %s" % (rep, time.time() - t0, st.tolist(),
                                                        [(r < 1e9 - 1).sum() for r in ranges]), flush=True)
    for i in range(n):
        print("gpu D1[%d] %s" % (i, digest(D1[i])))
    print("launches", jn.launch_count(), "total %.2fs" % (time.time() - t0))


if __name__ == "__main__":
    if "--make" in sys.argv:
        make()
    elif "--check" in sys.argv:
        check()
    else:
        run()
