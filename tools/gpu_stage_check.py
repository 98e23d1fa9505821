"""Stage-by-stage comparison of the CUDA path with the CPU oracle (run on a GPU box).

    python tools/gpu_stage_check.py [--full]

Prints, per configuration and per stage, the number of mismatching elements.  Test
infrastructure: loads oracle/ as the checker.
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol

jn = importlib.import_module("jackal-navigation_b200")
synth = importlib.import_module("jackal-navigation_b200.synth")

STAGES = ["desc1", "desc2", "dcan_raw", "dcan_incon", "dcan_final", "support", "tri1", "tri2", "planes1", "planes2",
          "grid1", "grid2", "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap",
          "D1_mean", "D2_mean", "D1", "D2"]


def compare(a, b):
    res = {}
    for k in STAGES:
        if k not in a or k not in b:
            continue
        x, y = a[k], b[k]
        if x.shape != y.shape:
            res[k] = "shape %s vs %s" % (x.shape, y.shape)
        else:
            n = int((x != y).sum())
            if n:
                res[k] = n
    return res


def scan_check(oracle, full):
    import scan_lib
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    cal = jn.Calibration(scan_lib.CALIB_YML)
    A = cal.arrays()
    for name, (W, H, dm, seed) in (("640x480", (640, 480, 64, 1)),) + ((("1920x1200_Kx3", (1920, 1200, 255, 1000)),) if full else ()):
        Q = np.array(fx["Q"][name])
        cal.set_q_matrix(Q)
        I1, I2, gt = synth.synth_pair(W, H, dm, seed)
        D1, _ = oracle.process(ol.robotics(dm), I1, I2)
        t = time.time(); gate_ref = sp.gate(Q, A["XR"], A["XT"], W, H); tg = time.time() - t
        sc = jn.ObstacleScan(cal, W, H)
        gate = sc.gate_cache()
        u8_ref = sp.convert_u8(D1)
        r_ref, m_ref = sp.scan(Q, A["XR"], A["XT"], gate_ref, u8_ref)
        r, m, u8 = sc.from_disparity(D1, want_u8=True)
        fin = r_ref < 1e9 - 1
        print("scan %s: gate mismatches %d (cpu gate %.1fs)  u8 mismatches %d  bins finite %d/%d same-set %s  max|dr| %.3e  n_points %d/%d" % (
            name, int((gate != gate_ref).sum()), tg, int((u8 != u8_ref).sum()), int(fin.sum()), m.n_finite,
            np.array_equal(fin, r < 1e9 - 1), float(np.abs(r[fin] - r_ref[fin]).max()) if fin.any() else 0.0, m.n_points, m_ref.n_points))
        print("   meta ours  %.12f %.12f %.9f %.9f" % (m.angle_min, m.angle_max, m.range_min, m.range_max))
        print("   meta port  %.12f %.12f %.9f %.9f" % (m_ref.angle_min, m_ref.angle_max, m_ref.range_min, m_ref.range_max))
        pts_ref = sp.points(Q, A["XR"], A["XT"], u8_ref)
        r2_ref, m2_ref = sp.scan_points(pts_ref)
        pts, r2, m2 = sc.points(D1)
        same = pts.shape == pts_ref.shape
        print("   -g path: points %d/%d  max|dp| %.3e  bins same-set %s max|dr| %.3e" % (
            len(pts), len(pts_ref), float(np.abs(pts - pts_ref).max()) if same and len(pts) else -1,
            np.array_equal(r2 < 1e9 - 1, r2_ref < 1e9 - 1),
            float(np.abs(r2[r2_ref < 1e9 - 1] - r2_ref[r2_ref < 1e9 - 1]).max()) if (r2_ref < 1e9 - 1).any() else 0.0))
        sc.close()


def main():
    full = "--full" in sys.argv
    oracle = ol.load("ref") or ol.load("port")
    print("oracle:", oracle.kind)
    cfgs = [(320, 240, 64, 1, {}), (333, 251, 100, 7, {}), (640, 480, 64, 1, {}),
            (640, 480, 255, 5, {"filter_median": 1, "postprocess_only_left": 0})]
    if full:
        cfgs += [(1920, 1200, 255, 1000, {}), (1920, 600, 255, 1001, {"filter_median": 1, "postprocess_only_left": 0})]
    summary = []
    for (W, H, dm, seed, kw) in cfgs:
        I1, I2, gt = synth.synth_pair(W, H, dm, seed)
        po = ol.robotics(dm, **kw)
        pj = jn.parameters(jn.ROBOTICS, disp_max=dm, **kw)
        a = oracle.stages(po, I1, I2)
        e = jn.Elas(pj)
        t = time.time()
        b = e.stages(I1, I2)
        dt = time.time() - t
        bad = compare(a, b)
        print("%dx%d dmax %d seed %d %s: rc %d/%d nsup %d/%d ntri %d/%d  %.3fs  mismatches: %s" % (
            W, H, dm, seed, kw, a["rc"], b["rc"], a["n_support"], b["n_support"], len(a["tri1"]), len(b["tri1"]), dt,
            bad if bad else "NONE"))
        # full process path
        D1 = np.zeros((H, W), np.float32); D2 = np.zeros((H, W), np.float32)
        rc = e.process(I1, I2, D1, D2, (W, H, W))
        d1bad = int((D1 != a["D1"]).sum()); d2bad = int((D2 != a["D2"]).sum())
        print("   process(): rc %d  D1 mismatches %d  D2 mismatches %d" % (rc, d1bad, d2bad))
        summary.append({"cfg": [W, H, dm, seed, kw], "stage_mismatch": {k: str(v) for k, v in bad.items()},
                        "process_mismatch": [d1bad, d2bad]})
        e.close()
    scan_check(oracle, full)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(summary, open(os.path.join(ROOT, "gpurun_out", "stage_check.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
