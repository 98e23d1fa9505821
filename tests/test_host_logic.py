"""CPU-only checks of the host side: the C ABI library loads and exports every declared
symbol, presets equal the reference's, the YAML reader, loud failure without a GPU."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest
import oracle_lib as ol
import scan_lib

ROOT = ol.ROOT


def declared_functions():
    names = set()
    for h in ("jn_elas.h", "jn_elas_debug.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"\b(jn_[a-z0-9_]+)\s*\(", txt):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol(jn):
    lib = jn.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libjn_elas.so does not export %s" % n


def test_presets_equal_reference_parameters(jn):
    """Elas::parameters(ROBOTICS / MIDDLEBURY), elas.h:87-144."""
    for setting, mk in ((jn.ROBOTICS, ol.robotics), (jn.MIDDLEBURY, ol.middlebury)):
        a, b = jn.parameters(setting), mk()
        for f, _ in ol.Params._fields_:
            assert getattr(a, f) == getattr(b, f), f


def test_presets_equal_compiled_reference(jn, ref):
    for setting in (0, 1):
        b = ol.Params()
        ref.lib.ref_params_default(C.byref(b), setting)
        a = jn.parameters(setting)
        for f, _ in ol.Params._fields_:
            assert getattr(a, f) == getattr(b, f), f


def test_calibration_yaml(jn):
    c = jn.Calibration(scan_lib.CALIB_YML).arrays()
    fx = scan_lib.fixtures()["calib"]
    for k in ("K1", "K2", "D1", "D2", "R", "T", "XR", "XT"):
        assert np.array_equal(c[k].reshape(-1), np.array(fx[k], np.float64).reshape(-1)), k
    with pytest.raises(jn.JnError):
        jn.Calibration("/nonexistent/calib.yml")


def test_calibration_yaml_missing_key(jn, tmp_path):
    p = tmp_path / "bad.yml"
    p.write_text("%YAML:1.0\nK1: !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [ 1., 0., 0., 0., 1., 0., 0., 0., 1. ]\n")
    with pytest.raises(jn.JnError):
        jn.Calibration(str(p))


def test_set_q(jn):
    c = jn.Calibration()
    c.set_q(338.27, 240.56, 679.54, -0.094)
    Q = c.arrays()["Q"]
    assert Q[0, 3] == -338.27 and Q[1, 3] == -240.56 and Q[2, 3] == 679.54 and Q[3, 2] == 1.0 / 0.094
    assert Q[3, 3] == 0 and Q[0, 0] == 1 and Q[1, 1] == 1


def test_no_cpu_fallback(jn):
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(jn.JnError, match="no CPU path"):
        jn.Elas(jn.parameters())


def test_rectifier_and_scan_have_no_cpu_fallback(jn):
    """The neighbours of the path (rectification, scan) also refuse to run without a device."""
    import torch
    import numpy as np
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = np.zeros((4, 4), np.float32)
    with pytest.raises(jn.JnError):
        jn.Rectifier(m, m)
    cal = jn.Calibration(os.path.join(ROOT, "tests", "golden", "calib_c920.yml"))
    cal.set_q(320.0, 240.0, 600.0, -0.1)
    with pytest.raises(jn.JnError):
        jn.ObstacleScan(cal, 64, 48)


def test_rectifier_rejects_bad_maps(jn):
    import numpy as np
    with pytest.raises(ValueError):
        jn.Rectifier(np.zeros((4, 4), np.float32), np.zeros((4, 5), np.float32))


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "jackal-navigation_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle/" not in txt and "oracle_lib" not in txt and "libelas_ref" not in txt, f


def test_prior_table_matches_port(jn, port):
    """P and plane_radius (elas.cpp:802-806) for both presets: ROBOTICS gives {-14,-9,-2}, radius 2."""
    p = ol.robotics()
    P = np.zeros(256, np.int32); r = C.c_int32(0)
    port.lib.port_prior(C.byref(p), P.ctypes.data_as(C.c_void_p), C.byref(r))
    assert r.value == 2 and list(P[:3]) == [-14, -9, -2]


def test_scan_compact_order(jn, port):
    ranges = np.full(90, 1e9)
    ranges[[3, 10, 80]] = [1.5, 2.5, 3.5]
    got = jn.scan_compact(ranges)
    assert list(got) == [3.5, 2.5, 1.5]            # k = 89..0, finite bins only
    assert list(scan_lib.ScanPort().compact(ranges)) == [3.5, 2.5, 1.5]


def test_synth_ground_truth(synth):
    I1, I2, gt = synth.synth_pair(320, 240, 64, 5)
    v, u = 200, 300
    d = gt[v, u]
    assert d == int(0.10 * 64 + 0.50 * 64 * v / 240)
    assert I2[v, u - d] == I1[v, u]
    assert gt[120, 160] == int(0.6 * 64)


def test_stereo_rectify_and_maps_match_opencv(jn):
    """csrc/calib.cu (OpenCV-free stereoRectify / initUndistortRectifyMap, point_cloud.cpp:543-554) against
    cv2 4.13 outputs (tests/golden/make_rectify_golden.py): the reference's own call on the shipped
    calibration at three image sizes, and perturbed calibrations incl. vertical stereo, no
    ZERO_DISPARITY and other alpha.  R1/R2/P1/P2/Q within 1e-9 relative, maps within 1e-3 px."""
    lib = jn.lib()
    P = C.c_void_p
    lib.jn_calib_stereo_rectify.argtypes = [C.POINTER(jn.Calib), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            P, P, P, P]
    lib.jn_calib_init_undistort_rectify_map.argtypes = [P, P, P, P, C.c_int, C.c_int, P, P]
    z = np.load(os.path.join(ROOT, "tests", "golden", "stereo_rectify_cv2.npz"))
    n = int(z["n"])
    assert n >= 10
    for i in range(n):
        g = lambda k: z["c%d_%s" % (i, k)]
        c = jn.Calib()
        for name, size in (("K1", 9), ("K2", 9), ("D1", 5), ("D2", 5), ("R", 9), ("T", 3)):
            v = np.asarray(g(name), np.float64).reshape(-1)
            assert v.size == size
            for j in range(size):
                getattr(c, name)[j] = float(v[j])
        cw, ch, nw, nh, zero = [int(x) for x in g("cfg")]
        R1 = np.zeros(9); R2 = np.zeros(9); P1 = np.zeros(12); P2 = np.zeros(12)
        ptr = lambda a: a.ctypes.data_as(P)
        rc = lib.jn_calib_stereo_rectify(C.byref(c), cw, ch, nw, nh, zero, float(g("alpha")), ptr(R1), ptr(R2), ptr(P1), ptr(P2))
        assert rc == 0 and c.has_q == 1
        for got, name in ((R1, "R1"), (R2, "R2"), (P1, "P1"), (P2, "P2"), (np.array(list(c.Q)), "Q")):
            ref = np.asarray(g(name), np.float64).reshape(-1)
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-10), (i, name, np.abs(got - ref).max())
        w, h = (nw, nh) if nw else (cw, ch)
        for K, D, Rr, Pp, kx, ky in ((g("K1"), g("D1"), R1, P1, "mx", "my"), (g("K2"), g("D2"), R2, P2, "mx2", "my2")):
            mx = np.zeros((h, w), np.float32); my = np.zeros((h, w), np.float32)
            K = np.ascontiguousarray(K, np.float64); D = np.ascontiguousarray(np.asarray(D, np.float64).reshape(-1))
            assert lib.jn_calib_init_undistort_rectify_map(ptr(K), ptr(D), ptr(Rr), ptr(Pp), w, h, ptr(mx), ptr(my)) == 0
            assert np.abs(mx[::23, ::29] - g(kx)).max() <= 1e-3 and np.abs(my[::23, ::29] - g(ky)).max() <= 1e-3, (i, kx)


def test_calibration_to_q_without_opencv(jn):
    """The whole init path of point_cloud.cpp:530-544 without OpenCV: YAML -> stereoRectify -> Q equals the
    cv2-generated fixture Q used everywhere else in the tests."""
    cal = jn.Calibration(scan_lib.CALIB_YML)
    fx = scan_lib.fixtures()["Q"]
    for name, (cw, ch, nw, nh, scale) in {"640x480": (640, 360, 640, 480, 1), "320x180": (640, 360, 320, 180, 1)}.items():
        cal.stereo_rectify(cw, ch, nw, nh)
        assert np.allclose(cal.arrays()["Q"], np.array(fx[name]), rtol=1e-9, atol=1e-10), name


def _vertexsort_lib(tmp_path):
    so = str(tmp_path / "libvs.so")
    subprocess.run(["g++", "-O1", "-std=c++14", "-shared", "-fPIC", "-I", os.path.join(ROOT, "jackal-navigation_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "vertexsort_host.cpp"), "-o", so], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.vs_sorted_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.vs_sorted_order_records.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return lib


def test_vertexsort_replay_keeps_the_copies_triangle_keeps(oracle, tmp_path):
    """csrc/vertexsort.cuh (run by one thread of the delaunay kernel when a point set holds coincident
    points) compiled for the host: the first entry of every run of equal points in its sorted order is
    the copy Triangle keeps (triangle.cpp:5446-5499, 6179-6195) = the vertex ids its triangles use."""
    lib = _vertexsort_lib(tmp_path)
    rng = np.random.default_rng(77)
    checked = 0
    for it in range(400):
        n = int(rng.integers(3, 400))
        mode = it % 4
        if mode == 0:
            pts = rng.integers(1, 7, size=(n, 2)) * 5                       # almost only copies
        elif mode == 1:
            pts = rng.integers(1, int(rng.integers(3, 40)), size=(n, 2)) * 5
        elif mode == 2:                                                     # right-image like: (u - d, v)
            pts = np.stack([rng.integers(1, 60, size=n) * 5 - rng.integers(0, 12, size=n) + 64,
                            rng.integers(1, 12, size=n) * 5], 1)
        else:
            pts = rng.integers(0, 2000, size=(n, 2))                        # (almost) no copies
        x = np.ascontiguousarray(pts[:, 0], np.int32); y = np.ascontiguousarray(pts[:, 1], np.int32)
        order = np.empty(n, np.int32)
        assert lib.vs_sorted_order(x.ctypes.data, y.ctypes.data, n, order.ctypes.data, 4096) == 0
        assert np.array_equal(np.sort(order), np.arange(n))
        order2 = np.empty(n, np.int32)     # (key, vertex) records instead of vertex numbers: same order
        assert lib.vs_sorted_order_records(x.ctypes.data, y.ctypes.data, n, order2.ctypes.data, 4096) == 0
        assert np.array_equal(order, order2)
        key = (x[order].astype(np.int64) << 13) | y[order]
        assert np.all(np.diff(key) >= 0)
        first = np.concatenate([[True], np.diff(key) != 0])
        survivors = np.sort(order[first])
        tri = oracle.triangulate(pts)
        if len(tri) == 0:
            continue                      # collinear / fewer than three distinct points: nothing to read the ids from
        assert np.array_equal(np.unique(tri), survivors), (it, mode, n)
        checked += 1 if len(survivors) < n else 0
    assert checked >= 150                 # sets with coincident points that produced triangles


def test_vertexsort_replay_reports_a_full_stack(tmp_path):
    lib = _vertexsort_lib(tmp_path)
    rng = np.random.default_rng(5)
    pts = rng.integers(0, 3000, size=(2000, 2)).astype(np.int32)
    x = np.ascontiguousarray(pts[:, 0]); y = np.ascontiguousarray(pts[:, 1])
    order = np.empty(2000, np.int32)
    assert lib.vs_sorted_order(x.ctypes.data, y.ctypes.data, 2000, order.ctypes.data, 2) == -1
    assert lib.vs_sorted_order(x.ctypes.data, y.ctypes.data, 2000, order.ctypes.data, 64) == 0


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("jn_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_bench_config_is_shared_by_both_arms_and_run_independent():
    """The driver compares the `config` object of the two arms: it must not contain anything arm- or run-specific."""
    import argparse
    b = _bench_module()
    a = argparse.Namespace(width=1920, height=1200, disp_max=255, batch=64, scene="random_dot", config="robotics",
                           total_frames=0)
    c1, c2 = b.bench_config(a, 1), b.bench_config(a, 1)
    assert c1 == c2 and c1["workload"].startswith("1920x1200 disp_max=255 ROBOTICS")
    assert set(c1) == {"workload", "frames_per_step_per_gpu", "pairs", "l2", "parallelism", "total_frames"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(') == 2          # GPU arm and reference arm
    a.total_frames = 1024
    assert "1000..2023" in b.bench_config(a, 8)["pairs"]     # BASELINE config C4


def test_bench_stage_roofline_figures():
    """179 N bytes per frame end to end (ROBOTICS), 251 N for C5; the support matcher's SAD count
    against a brute-force enumeration of elas.cpp:269-373's loops."""
    b = _bench_module()
    assert abs(sum(b.STAGE_BYTES_N.values()) + 4 - 179) < 0.2
    assert abs(sum(b.STAGE_BYTES_N_C5.values()) + 4 - 251) < 0.2
    W, H, dm = 640, 480, 64
    n = 0
    for vc in range(1, -(-H // 5)):
        v = vc * 5
        if v < 5 or v > H - 6:
            continue
        for uc in range(1, -(-W // 5)):
            u = uc * 5
            if u < 5 or u > W - 6:
                continue
            for right in (0, 1):
                d_hi = min(dm, (W - u - 5) if right else (u - 5))
                if d_hi - 0 < 10:                        # elas.cpp:320-331: ranges shorter than 10 are rejected
                    continue
                n += (d_hi + 1) * 16                     # 4 blocks of 16 bytes = 16 four-byte SADs per disparity
    assert abs(b.support_warp_sads(W, H, dm) - n / 32.0) < 1e-6 * n
    st = {"descriptor": 1.3, "support": 4.8, "delaunay": 0.7, "planes_grid": 0.4, "raster": 1.2, "dense_match": 2.515,
          "post": 3.0}
    r = b.stage_roofline(st, 1920, 1200, 255, 64, 6447.8, 1965.0)
    assert abs(r["dense_match"]["frac"] - 72 * 1920 * 1200 * 64 / 6447.8e9 / 2.515e-3) < 1e-9
    assert r["support"]["bound"] == "integer pipe" and 0.3 < r["support"]["frac"] < 1.0
    assert r["delaunay"]["frac"] is None and r["raster"]["frac"] is None
    assert b.stage_roofline(st, 1920, 1200, 255, 64, 6447.8, None)["support"]["frac"] == r["support"]["frac"]


def test_documents_name_only_declared_entry_points():
    """INTEGRATION.md / DESIGN.md / README.md show a maintainer what to bind: every jn_* name they mention is
    declared in include/*.h (a prefix such as jn_navigate_ or jn_scan_ stands for a family of declared names)."""
    decl = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        decl |= set(re.findall(r"\b(jn_[a-z0-9_]+)\b", open(os.path.join(ROOT, "include", h)).read()))
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        names = set(re.findall(r"\b(jn_[a-z0-9_]+)\b", open(os.path.join(ROOT, doc)).read()))
        assert names, doc
        missing = sorted(n for n in names if n not in decl and not any(d.startswith(n) for d in decl))
        assert not missing, "%s mentions undeclared %s" % (doc, missing)


def test_python_wrappers_reject_undersized_arrays_before_the_c_call(jn):
    """The C side reads W*H elements through raw pointers: the Python mirror refuses arrays that are too small
    (ValueError) instead of letting the library read past them.  No device needed: the check comes first, and a
    right-sized call on a NULL handle comes back as the library's JN_ERR_ARG."""
    class Fake:
        W, H, _h = 64, 48, None
    small = np.zeros((10, 10), np.float32)
    full = np.zeros((48, 64), np.float32)
    for call in (lambda: jn.ObstacleScan.from_disparity(Fake, small),
                 lambda: jn.ObstacleScan.points(Fake, small),
                 lambda: jn.ObstacleScan.pointcloud(Fake, small, np.zeros((48, 64), np.uint8)),
                 lambda: jn.ObstacleScan.pointcloud(Fake, full, np.zeros((20, 64), np.uint8)),
                 lambda: jn.ObstacleScan.pointcloud(Fake, full, np.zeros((48, 64, 4), np.uint8)),
                 lambda: jn.scan_compact(np.zeros(10))):
        with pytest.raises(ValueError):
            call()
    for call in (lambda: jn.ObstacleScan.from_disparity(Fake, full),
                 lambda: jn.ObstacleScan.points(Fake, full),
                 lambda: jn.ObstacleScan.pointcloud(Fake, full, np.zeros((48, 64, 3), np.uint8))):
        with pytest.raises(jn.JnError):
            call()
    assert len(jn.scan_compact(np.full(90, 1e9))) == 0
