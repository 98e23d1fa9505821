"""Pins the plain-C oracle port (oracle/elas_port.c, delaunay_port.c) to the reference:
against the committed golden vectors (generated from the unmodified reference, see
tests/golden/make_elas_golden.py) and, where oracle/_ref is built, against the compiled
reference itself.  CPU only."""
import numpy as np
import pytest
import golden_util as gu
import oracle_lib as ol

STAGE_KEYS = gu.EXACT + ["desc1", "desc2", "grid1", "grid2"]


@pytest.mark.parametrize("name,H", [("elas_robotics_160x120.npz", 120), ("elas_c5_200x150.npz", 150),
                                    ("elas_sub_240x180.npz", 180)])
def test_port_matches_golden(port, name, H):
    z, p = gu.load(name)
    o = port.stages(p, z["I1"], z["I2"])
    assert o["rc"] == 0
    gu.check(o, z, H)


def test_port_delaunay_matches_golden_triangle_output(port):
    z = np.load(gu.GOLD + "/delaunay_cases.npz")
    names = [k[4:] for k in z.files if k.startswith("pts_")]
    assert len(names) >= 8
    for n in names:
        got = port.triangulate(z["pts_" + n])
        assert np.array_equal(got, z["tri_" + n]), n


@pytest.mark.parametrize("W,H,dm,seed,kw", [
    (320, 240, 64, 1, {}),
    (333, 251, 100, 7, {}),                                           # width not a multiple of 16
    (640, 480, 64, 1, {}),                                            # BASELINE config C1
    (320, 240, 64, 3, {"filter_median": 1, "postprocess_only_left": 0}),
    (256, 192, 255, 9, {"ipol_gap_width": 7, "speckle_size": 50, "lr_threshold": 1}),
    (320, 240, 64, 1, {"subsampling": 1}),                            # half-resolution maps
    (333, 251, 100, 7, {"subsampling": 1, "filter_median": 1, "postprocess_only_left": 0}),
    (326, 241, 80, 6, {"subsampling": 1, "candidate_stepsize": 4, "add_corners": 1}),
])
def test_port_matches_compiled_reference_every_stage(ref, port, synth, W, H, dm, seed, kw):
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    p = ol.robotics(dm, **kw)
    a, b = ref.stages(p, I1, I2), port.stages(p, I1, I2)
    assert a["rc"] == b["rc"] == 0
    for k in STAGE_KEYS:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), "%s: %d mismatches" % (k, int((a[k] != b[k]).sum()))


@pytest.mark.parametrize("W,H,dm,seed,kw", [
    (640, 480, 64, 1, {}),
    (333, 251, 100, 7, {"filter_median": 1, "postprocess_only_left": 0}),
    (480, 360, 128, 4, {"subsampling": 1}),
])
def test_port_matches_reference_on_textured_scene(ref, port, synth, W, H, dm, seed, kw):
    """Second scene family (1/f texture, sub-pixel disparities slanted in u and v, occluding boxes,
    textureless patch): more, less regular support points and real matching failures."""
    I1, I2, gt = synth.textured_pair(W, H, dm, seed)
    p = ol.robotics(dm, **kw)
    a, b = ref.stages(p, I1, I2), port.stages(p, I1, I2)
    assert a["rc"] == b["rc"] == 0
    for k in STAGE_KEYS:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), "%s: %d mismatches" % (k, int((a[k] != b[k]).sum()))
    if W >= 640:
        valid = a["D1"] >= 0
        assert valid.mean() > 0.8 and (np.abs(a["D1"] - gt)[valid] <= 1).mean() > 0.97


@pytest.mark.parametrize("scene,W,H,dm,seed,kw", [
    ("random_dot", 320, 240, 64, 4, {"sradius": 4.0}),                                   # plane radius 4
    ("random_dot", 333, 251, 100, 6, {"sigma": 2.0, "sradius": 3.5, "match_texture": 3}),  # plane radius 7
    ("random_dot", 320, 240, 64, 4, {"sradius": 5.0, "subsampling": 1}),
    ("textured", 640, 480, 64, 8, {"lr_threshold": 5, "incon_threshold": 8}),            # coincident right-image points
    ("textured", 640, 480, 64, 3, {"candidate_stepsize": 2, "lr_threshold": 4, "incon_threshold": 8}),
])
def test_port_matches_reference_off_preset_parameters(ref, port, synth, scene, W, H, dm, seed, kw):
    """The parameter sets of the GPU tests that leave the presets: larger plane radii (prior table, plane range)
    and cross-check tolerances that let two support points share a right-image position, where the copy Triangle
    keeps depends on its seeded quicksort (reproduced by oracle/delaunay_port.c)."""
    I1, I2, _ = synth.SCENES[scene](W, H, dm, seed)
    p = ol.robotics(dm, **kw)
    a, b = ref.stages(p, I1, I2), port.stages(p, I1, I2)
    assert a["rc"] == b["rc"] == 0
    for k in STAGE_KEYS:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), "%s: %d mismatches" % (k, int((a[k] != b[k]).sum()))


def test_port_middlebury_matches_reference(ref, port, synth):
    I1, I2, _ = synth.synth_pair(320, 240, 64, 3)
    p = ol.middlebury(64)
    a, b = ref.stages(p, I1, I2), port.stages(p, I1, I2)
    for k in STAGE_KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_port_delaunay_matches_triangle_random(ref, port):
    rng = np.random.default_rng(0)
    for it in range(300):
        n = int(rng.integers(3, 300))
        mode = it % 5
        if mode == 0:
            pts = rng.integers(0, rng.integers(2, 30), size=(n, 2)) * 5
        elif mode == 1:
            pts = rng.integers(0, 2000, size=(n, 2))
        elif mode == 2:
            pts = rng.integers(0, 6, size=(n, 2)) * 5          # many duplicates
        elif mode == 3:
            pts = np.stack([rng.integers(0, 50, size=n) * 5, np.full(n, 10)], 1)   # collinear
        else:
            u = rng.integers(1, 380, size=n) * 5; v = rng.integers(1, 240, size=n) * 5
            pts = np.stack([u - rng.integers(0, 60, size=n), v], 1)
        if mode in (0, 1, 4):
            pts = np.unique(pts, axis=0)
            rng.shuffle(pts)
        if len(pts) < 3:
            continue
        assert np.array_equal(ref.triangulate(pts), port.triangulate(pts)), (it, mode)


def test_few_support_points_leave_outputs_untouched(port):
    """elas.cpp:66-71: a textureless pair has no support point; D1/D2 keep the caller's values."""
    I = np.full((120, 160), 77, np.uint8)
    p = ol.robotics(32)
    D1 = np.full((120, 160), 5.0, np.float32); D2 = D1.copy()
    import ctypes as C
    dims = (C.c_int32 * 3)(160, 120, 160)
    f = port.fn("elas_process"); f.restype = C.c_int
    rc = f(C.byref(p), I.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p),
           D1.ctypes.data_as(C.c_void_p), D2.ctypes.data_as(C.c_void_p), dims)
    assert rc == 1 and (D1 == 5.0).all() and (D2 == 5.0).all()


def test_adaptive_mean_step_weights(port):
    """the reference's mask turns the bilateral weight into a step function:
    4 for |delta| < 2, 2 for 2 <= |delta| < 8, 0 for |delta| >= 8 (elas.cpp:1320, 1411-1436)."""
    import ctypes as C
    H, W = 16, 32

    def weight(delta):
        a = abs(delta)
        return 4.0 if a < 2 else (2.0 if a < 8 else 0.0)

    for delta in (0.5, 1.0, 1.99, 2.0, 3.0, 7.9, 8.0, 10.0, 31.0, 32.0, 60.0, -1.5, -7.0):
        D = np.full((H, W), 20.0, np.float32)
        D[:, 16] = 20.0 + delta                       # one outlier column -> vertical pass is neutral
        ref_row = D[8].copy()
        port.lib.port_adaptive_mean(W, H, D.ctypes.data_as(C.c_void_p))
        # centre 14 sees columns 10..17 (offsets -4..+3): seven samples of 20 and the outlier
        w = weight(delta)
        expect = np.float32((7 * 4.0 * 20.0 + w * (20.0 + delta)) / (7 * 4.0 + w))
        assert abs(D[8, 14] - expect) < 1e-4, (delta, D[8, 14], expect)
        assert D[8, 5] == 20.0 and ref_row[5] == 20.0


def test_remap_port_matches_opencv_golden(port):
    """oracle/remap_port.c against outputs of cv2.remap itself (calibration maps of the shipped YAML and
    adversarial maps: integer / tie / negative / far-outside coordinates)."""
    n = 0
    for name, src, mx, my, dst in ol.remap_golden_cases():
        got = ol.port_remap(src, mx, my)
        assert np.array_equal(got, dst), (name, int((got != dst).sum()))
        roi = (3, 2, mx.shape[1] - 7, mx.shape[0] - 5)
        assert np.array_equal(ol.port_remap(src, mx, my, roi), dst[2:2 + roi[3], 3:3 + roi[2]]), name
        n += 1
    assert n == 3


def test_scan_port_matches_opencv_arithmetic():
    """The parts of point_cloud.cpp's per-pixel arithmetic that live in OpenCV (cv::Mat products = cv::gemm,
    convertTo(CV_8U) = saturate_cast), evaluated by cv2 4.13 itself (tests/golden/make_scan_golden.py)."""
    import scan_lib
    z = np.load(gu.GOLD + "/scan_cv2.npz")
    sp = scan_lib.ScanPort()
    ox, oy = int(z["ox"]), int(z["oy"])
    assert np.array_equal(sp.points(z["Q"], z["XR"], z["XT"], z["dmap"], ox, oy), z["pts"])
    sub = np.zeros_like(z["dmap"]); sub[::3, ::3] = z["dmap"][::3, ::3]
    assert np.array_equal(sp.points(z["Qg"], z["XR"], z["XT"], sub, ox, oy), z["ptsg"])      # dense Q
    assert np.array_equal(sp.convert_u8(z["D"]), z["u8"])


def test_scan_port_matches_statement_by_statement_execution():
    """oracle/scan_port.c against point_cloud.cpp:104-147, 149-211, 213-296, 321-349 executed statement by
    statement in Python with cv2.gemm doing the cv::Mat products (tests/golden/make_scan_statement_golden.py):
    gate cache (incl. the 256 -> 0 wrap of H8), u8 conversion, 90-bin scan, extrema, compaction, the -g point
    list and the scan from the points -- all bit-exact (same libm)."""
    import scan_lib
    z = np.load(gu.GOLD + "/scan_statements.npz")
    sp = scan_lib.ScanPort()
    XR, XT = z["XR"], z["XT"]
    wrapped = 0
    for name in z["names"]:
        g = lambda k: z["%s_%s" % (name, k)]
        W, H, ox, oy = [int(x) for x in g("dims")]
        Q = g("Q")
        gate = sp.gate(Q, XR, XT, W, H, ox, oy)
        assert np.array_equal(gate, g("gate")), name
        wrapped += int((gate[..., 0] == 0).sum())
        u8 = sp.convert_u8(g("D"))
        assert np.array_equal(u8, g("u8")), name
        r, m = sp.scan(Q, XR, XT, gate, u8, ox, oy)
        assert np.array_equal(r, g("scan")), name
        meta = g("meta")
        assert (m.angle_min, m.angle_max, m.range_min, m.range_max, m.n_points) == tuple(meta[:4]) + (int(meta[4]),), name
        assert np.array_equal(sp.compact(r), g("compact")), name
        pts = sp.points(Q, XR, XT, u8, ox, oy)
        gp = g("pts")
        assert pts.shape == gp.shape and np.array_equal(pts, gp, equal_nan=True), name
        r2, m2 = sp.scan_points(pts)
        assert np.array_equal(r2, g("scan_p")), name
        mp_ = g("meta_p")
        assert (m2.angle_min, m2.angle_max, m2.range_min, m2.range_max) == tuple(mp_[:4]), name
    assert wrapped > 0          # the wrap case is in the fixture


def test_raster_overlaps_stay_on_span_borders(port, synth):
    """Design invariant behind raster_kernel's plain stores: where two scan-converted triangles of
    computeDisparity (elas.cpp:843-903) cover the same pixel, the pixel is in the first or last row of
    the column span of BOTH -- never strictly inside a span.  Pipeline triangulations (lattice support
    points, left and right image) and adversarial random point sets."""
    import ctypes as C
    f = port.lib.port_raster_overlap_study
    P = C.c_void_p

    def study(W, H, sup3, tri, right):
        out = (C.c_int64 * 4)()
        f(W, H, np.ascontiguousarray(sup3, np.int32).ctypes.data_as(P),
          np.ascontiguousarray(tri, np.int32).ctypes.data_as(P), len(tri), right, out)
        return list(out)

    multi = 0
    for W, H, dm, seed, preset in ((320, 240, 64, 1, ol.robotics), (640, 480, 255, 21, ol.robotics),
                                   (320, 240, 64, 4, ol.middlebury)):
        I1, I2, _ = synth.synth_pair(W, H, dm, seed)
        st = port.stages(preset(dm), I1, I2, want_desc=False, want_grid=False)
        for right in (0, 1):
            o = study(W, H, st["support"], st["tri2" if right else "tri1"], right)
            assert o[2] == 0, (W, H, seed, right, o)
            multi += o[1]
    rng = np.random.default_rng(123)
    for it in range(300):
        W, H, n = int(rng.integers(40, 400)), int(rng.integers(30, 300)), int(rng.integers(3, 300))
        if it % 3 == 0:
            u, v = rng.integers(1, max(2, W // 5), n) * 5, rng.integers(1, max(2, H // 5), n) * 5
        elif it % 3 == 1:
            u, v = rng.integers(0, W, n), rng.integers(0, H, n)
        else:
            u, v = rng.integers(0, 12, n) * 3, rng.integers(0, 12, n) * 3
        sup = np.unique(np.stack([u, v], 1), axis=0)
        if len(sup) < 3:
            continue
        sup3 = np.concatenate([sup, rng.integers(0, 40, (len(sup), 1))], 1).astype(np.int32)
        for right in (0, 1):
            pts = sup3[:, :2].copy()
            s_in = sup3.copy()
            if right:
                pts[:, 0] = sup3[:, 0] - sup3[:, 2] + 64
                s_in[:, 0] += 64
            tri = port.triangulate(np.ascontiguousarray(pts, np.int32))
            if len(tri):
                o = study(W + 128, H, s_in, tri, right)
                assert o[2] == 0, (it, right, o)
                multi += o[1]
    assert multi > 0          # overlaps do occur (on span borders): the atomics there are needed
    # the largest size the plain-store path is used for (raster_kernel: W, H <= 2048), with few points
    # = long, thin triangles and steep hull edges, where the float edge evaluation is least accurate
    for it in range(6):
        W = H = 2048
        n = (12, 60, 400)[it % 3]
        u, v = rng.integers(0, W, n), rng.integers(0, H, n)
        if it >= 3:
            u[: n // 3] = rng.integers(0, 4, n // 3) + (0, W // 2, W - 4)[it - 3]    # near-vertical edges
        sup = np.unique(np.stack([u, v], 1), axis=0)
        sup3 = np.concatenate([sup, np.zeros((len(sup), 1), np.int64)], 1).astype(np.int32)
        tri = port.triangulate(np.ascontiguousarray(sup3[:, :2]))
        o = study(W, H, sup3, tri, 0)
        assert o[2] == 0, ("2048", it, o)
