"""Parity of the CUDA path with the oracle, through the C ABI (run with -m gpu on a B200).

Integer stages (descriptors, support points, triangles, grid, SAD argmin disparities, L/R,
segments) must be bit-exact; with identical arithmetic order and no FMA the float stages
(planes, gap interpolation, adaptive mean, median) are bit-exact too, so every comparison
below is exact unless a tolerance is written next to it."""
import os
import numpy as np
import pytest
import golden_util as gu
import oracle_lib as ol
import scan_lib

pytestmark = pytest.mark.gpu

STAGES = ["desc1", "desc2", "dcan_raw", "dcan_incon", "dcan_final", "support", "tri1", "tri2", "planes1", "planes2",
          "grid1", "grid2", "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap",
          "D1_mean", "D2_mean", "D1", "D2"]


def assert_stages_equal(a, b, keys=STAGES):
    assert a["rc"] == b["rc"]
    for k in keys:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, (k, x.shape, y.shape)
        assert np.array_equal(x, y), "%s: %d of %d elements differ" % (k, int((x != y).sum()), x.size)


@pytest.mark.parametrize("name,H", [("elas_robotics_160x120.npz", 120), ("elas_c5_200x150.npz", 150),
                                    ("elas_sub_240x180.npz", 180)])
def test_cuda_matches_golden(jn, name, H):
    z, p = gu.load(name)
    pj = jn.parameters.from_buffer_copy(bytes(p))
    e = jn.Elas(pj)
    o = e.stages(z["I1"], z["I2"])
    assert o["rc"] == 0
    gu.check(o, z, H)
    e.close()


@pytest.mark.parametrize("W,H,dm,seed,kw", [
    (320, 240, 64, 1, {}),
    (333, 251, 100, 7, {}),                                            # ragged width/height
    (640, 480, 64, 1, {}),                                             # BASELINE C1
    (640, 480, 255, 5, {"filter_median": 1, "postprocess_only_left": 0}),
    (256, 192, 255, 9, {"ipol_gap_width": 7, "speckle_size": 50, "lr_threshold": 1}),
    (320, 240, 64, 3, {"filter_adaptive_mean": 0}),
    (320, 240, 64, 8, {"ipol_gap_width": 40, "postprocess_only_left": 0}),          # CTA-scan gap kernels
    (320, 240, 64, 8, {"ipol_gap_width": 5000, "filter_median": 1, "filter_adaptive_mean": 0,
                       "postprocess_only_left": 0, "match_texture": 0, "gamma": 5.0, "sradius": 3.0,
                       "support_threshold": 0.95}),                                    # MIDDLEBURY minus add_corners
    (320, 240, 64, 4, {"sradius": 4.0}),                               # plane radius 4: run-time radius kernel
    (333, 251, 100, 6, {"sigma": 2.0, "sradius": 3.5, "match_texture": 3}),   # plane radius 7 (the largest covered)
    (2200, 1300, 255, 2, {}),                                          # lattice too large for the smem filter
    (1920, 600, 255, 1001, {}),                                        # C3 with -h 600 crop
])
def test_every_stage_matches_oracle(jn, oracle, synth, W, H, dm, seed, kw):
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    a = oracle.stages(ol.robotics(dm, **kw), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm, **kw))
    b = e.stages(I1, I2)
    assert_stages_equal(a, b)
    e.close()


@pytest.mark.parametrize("seed,kw", [
    (8, {"lr_threshold": 5, "incon_threshold": 8}),
    (3, {"candidate_stepsize": 2, "lr_threshold": 4, "incon_threshold": 8}),
])
def test_coincident_right_image_points_follow_triangle(jn, oracle, synth, seed, kw):
    """With a cross-check tolerance of at least half the candidate step, two support points of one row
    can share their right-image position (u - d, v).  Triangle drops all but one copy -- the one its
    randomised quicksort puts first (triangle.cpp:5446-5499, 6179-6195) -- and the copy decides the
    vertex ids and planes of the right-image triangles around it.  On these pairs the survivor is NOT
    the lowest support index; the kernel replays the quicksort (csrc/vertexsort.cuh)."""
    W, H, dm = 640, 480, 64
    I1, I2, _ = synth.textured_pair(W, H, dm, seed)
    a = oracle.stages(ol.robotics(dm, **kw), I1, I2)
    sup = np.asarray(a["support"])
    right = np.stack([sup[:, 0] - sup[:, 2], sup[:, 1]], 1)
    assert len(np.unique(right, axis=0)) < len(right), "the pair no longer produces coincident points"
    _, lowest = np.unique((right[:, 0].astype(np.int64) << 16) | right[:, 1], return_index=True)
    assert not np.array_equal(np.sort(lowest), np.unique(a["tri2"])), "lowest index = Triangle's choice here"
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm, **kw))
    try:
        # (-1,-1): ordering in shared memory; (0,-1): the path of very large point sets (occupancy-grid
        # ranking in global memory), which picks its survivors the same way
        for limits in ((-1, -1), (0, -1)):
            jn.lib().jn_debug_delaunay_limits(*limits)
            b = e.stages(I1, I2)
            assert_stages_equal(a, b)
    finally:
        jn.lib().jn_debug_delaunay_limits(-1, -1)
    e.close()


@pytest.mark.parametrize("W,H,dm,seed", [(320, 240, 64, 3), (640, 480, 255, 12), (333, 251, 100, 7)])
def test_middlebury_preset_matches_oracle(jn, oracle, synth, W, H, dm, seed):
    """Elas::parameters(MIDDLEBURY): add_corners (4 corner support points + 2 shifted copies,
    border extrapolation in the gap interpolation), median, both images, plane radius 3."""
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    a = oracle.stages(ol.middlebury(dm), I1, I2)
    e = jn.Elas(jn.parameters(jn.MIDDLEBURY, disp_max=dm))
    b = e.stages(I1, I2)
    assert_stages_equal(a, b)
    assert (b["support"][-6:, :2] == [[0, 0], [0, H - 1], [W - 1, 0], [W - 1, H - 1],
                                       [W - 1 + b["support"][-2, 2], 0], [W - 1 + b["support"][-1, 2], H - 1]]).all()
    e.close()


def test_add_corners_on_textureless_pair(jn, oracle):
    """No support point at all: the corners get disparity 0 and the pipeline continues (6 points)."""
    I = np.full((120, 160), 77, np.uint8)
    a = oracle.stages(ol.middlebury(32), I, I)
    e = jn.Elas(jn.parameters(jn.MIDDLEBURY, disp_max=32))
    b = e.stages(I, I)
    assert a["rc"] == b["rc"] == 0 and b["n_support"] == 6
    # With disparity 0 the shifted corner copies coincide with the right corners.  Triangle keeps
    # whichever copy its randomised quicksort puts first; the kernel replays that quicksort
    # (csrc/vertexsort.cuh), so the vertex ids of the triangles are the reference's as well.
    assert_stages_equal(a, b, STAGES)
    e.close()


def test_full_size_1920x1200_matches_oracle(jn, oracle, synth):
    """BASELINE config C3 / the bench workload, ROBOTICS and C5 (median + both images)."""
    W, H, dm = 1920, 1200, 255
    I1, I2, gt = synth.synth_pair(W, H, dm, 1000)
    for kw in ({}, {"filter_median": 1, "postprocess_only_left": 0}):
        a = oracle.stages(ol.robotics(dm, **kw), I1, I2, want_desc=False, want_grid=False)
        e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm, **kw))
        b = e.stages(I1, I2, want_desc=False, want_grid=False)
        assert_stages_equal(a, b, [k for k in STAGES if not k.startswith(("desc", "grid"))])
        valid = b["D1"] >= 0
        assert valid.mean() > 0.85
        assert (np.abs(b["D1"] - gt)[valid] <= 1).mean() > 0.999
        e.close()


@pytest.mark.parametrize("W,H,dm,seed,kw", [
    (640, 480, 64, 1, {}),
    (333, 251, 100, 7, {"filter_median": 1, "postprocess_only_left": 0}),
    (480, 360, 128, 4, {"subsampling": 1}),
    (1920, 1200, 255, 1000, {}),                       # > 9 000 support points: the large Delaunay path
    (1920, 1200, 255, 1001, {"filter_median": 1, "postprocess_only_left": 0}),
])
def test_textured_scene_matches_oracle(jn, oracle, synth, W, H, dm, seed, kw):
    """Second scene family: 1/f texture, sub-pixel disparities slanted in u and v, occluding boxes, a
    textureless patch.  Every stage bit-exact, as on the random-dot scenes."""
    I1, I2, gt = synth.textured_pair(W, H, dm, seed)
    big = W * H > 10 ** 6
    a = oracle.stages(ol.robotics(dm, **kw), I1, I2, want_desc=not big, want_grid=not big)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm, **kw))
    b = e.stages(I1, I2, want_desc=not big, want_grid=not big)
    assert_stages_equal(a, b, [k for k in STAGES if not (big and k.startswith(("desc", "grid")))])
    if big:
        assert b["n_support"] > 8192
    e.close()


def test_delaunay_large_point_set_paths(jn, oracle, synth):
    """The Delaunay kernel has shared-memory fast paths (bitonic sort, u16 tables) and
    global-memory paths for large point sets; force the latter and compare again."""
    W, H, dm = 640, 480, 255
    I1, I2, _ = synth.synth_pair(W, H, dm, 5)
    a = oracle.stages(ol.robotics(dm), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    try:
        for limits in ((0, 0), (-1, 0), (0, -1)):
            jn.lib().jn_debug_delaunay_limits(*limits)
            b = e.stages(I1, I2)
            assert_stages_equal(a, b, ["support", "tri1", "tri2", "planes1", "planes2", "D1", "D2"])
    finally:
        jn.lib().jn_debug_delaunay_limits(-1, -1)
    e.close()


def test_dense_grid_list_overflow_path(jn, oracle, synth):
    """The dense matcher reads a compact candidate list per grid cell and falls back to the
    bit set for cells with more than 16 candidates; force the fallback and compare again."""
    W, H, dm = 640, 480, 255
    I1, I2, _ = synth.synth_pair(W, H, dm, 5)
    a = oracle.stages(ol.robotics(dm), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    try:
        for limit in (0, 3):
            jn.lib().jn_debug_grid_list_limit(limit)
            b = e.stages(I1, I2)
            assert_stages_equal(a, b, ["grid1", "grid2", "D1_raw", "D2_raw", "D1", "D2"])
    finally:
        jn.lib().jn_debug_grid_list_limit(-1)
    e.close()


def test_process_is_a_drop_in(jn, oracle, synth):
    """Elas::process semantics: same maps, caller-owned buffers, bytes_per_line honoured."""
    W, H, dm = 333, 251, 64
    I1, I2, _ = synth.synth_pair(W, H, dm, 21)
    R1, R2 = oracle.process(ol.robotics(dm), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    D1 = np.zeros((H, W), np.float32); D2 = np.zeros((H, W), np.float32)
    assert e.process(I1, I2, D1, D2, (W, H, W)) == 0
    assert np.array_equal(D1, R1) and np.array_equal(D2, R2)
    # padded rows (bytes_per_line > width), elas.cpp:44-52
    bpl = 352
    P1 = np.full((H, bpl), 255, np.uint8); P2 = np.full((H, bpl), 255, np.uint8)
    P1[:, :W] = I1; P2[:, :W] = I2
    D1b = np.zeros((H, W), np.float32); D2b = np.zeros((H, W), np.float32)
    assert e.process(P1, P2, D1b, D2b, (W, H, bpl)) == 0
    assert np.array_equal(D1b, R1) and np.array_equal(D2b, R2)
    e.close()


def test_few_support_points_leave_outputs_untouched(jn, capsys):
    """elas.cpp:66-71."""
    I = np.full((120, 160), 77, np.uint8)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=32))
    D1 = np.full((120, 160), 5.0, np.float32); D2 = D1.copy()
    assert e.process(I, I, D1, D2, (160, 120, 160)) == jn.JN_FEW_SUPPORT
    assert (D1 == 5.0).all() and (D2 == 5.0).all()
    assert "Need at least 3 support points" in capsys.readouterr().out
    e.close()


@pytest.mark.parametrize("W,H,dm,seed,preset,kw", [
    (320, 240, 64, 1, "robotics", {}),
    (333, 251, 100, 7, "robotics", {}),                                   # odd width and height
    (640, 480, 255, 5, "robotics", {"filter_median": 1, "postprocess_only_left": 0}),
    (326, 241, 80, 6, "robotics", {"candidate_stepsize": 4}),             # even step stays as it is
    (320, 240, 64, 8, "robotics", {"ipol_gap_width": 40, "speckle_size": 30}),
    (320, 240, 64, 3, "middlebury", {}),
    (640, 480, 255, 12, "middlebury", {}),
    (320, 240, 64, 4, "robotics", {"sradius": 5.0}),                      # run-time plane radius, subsampled
    (1920, 1200, 255, 1001, "robotics", {}),
])
def test_subsampling_every_stage_matches_oracle(jn, oracle, synth, W, H, dm, seed, preset, kw):
    """param.subsampling = 1: descriptors on even rows only, lattice step 6, even pixels matched into
    (H/2) x (W/2) maps, post-processing at half resolution (elas.cpp:380, 693, 877-896, 914-1499)."""
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    a = oracle.stages(getattr(ol, preset)(dm, subsampling=1, **kw), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS if preset == "robotics" else jn.MIDDLEBURY, disp_max=dm, subsampling=1, **kw))
    b = e.stages(I1, I2)
    assert b["D1"].shape == (H // 2, W // 2)
    assert_stages_equal(a, b)
    D1 = np.full((H // 2, W // 2), 3.0, np.float32); D2 = D1.copy()
    assert e.process(I1, I2, D1, D2, (W, H, W)) == jn.JN_OK
    assert np.array_equal(D1, a["D1"]) and np.array_equal(D2, a["D2"])
    e.close()


def test_unsupported_parameters_fail_loudly(jn):
    I = np.zeros((120, 160), np.uint8); D = np.zeros((120, 160), np.float32)
    for kw in ({"sigma": 3.0, "sradius": 3.0},):       # plane radius 9 > 7
        e = jn.Elas(jn.parameters(jn.ROBOTICS, **kw))
        with pytest.raises(jn.JnError):
            e.process(I, I, D, D.copy(), (160, 120, 160))
        e.close()


def test_batch_equals_single_frames(jn, oracle, synth):
    """Frames of a batch are independent: a batch (with one textureless frame in the middle)
    gives exactly the per-frame results, in order."""
    import torch
    W, H, dm, B = 320, 240, 64, 5
    L, R = synth.synth_batch(W, H, dm, [31, 32, 33, 34, 35])
    L[2] = 9; R[2] = 9                      # frame 2 has no support points
    dev = torch.device("cuda", 0)
    dL = torch.from_numpy(L).to(dev); dR = torch.from_numpy(R).to(dev)
    dD1 = torch.full((B, H, W), 7.0, dtype=torch.float32, device=dev)
    dD2 = torch.full((B, H, W), 7.0, dtype=torch.float32, device=dev)
    st = torch.full((B,), -9, dtype=torch.int32, device=dev)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    s = torch.cuda.Stream()
    e.process_batch(dL.data_ptr(), dR.data_ptr(), dD1.data_ptr(), dD2.data_ptr(), st.data_ptr(), (W, H, W), B,
                    s.cuda_stream)
    torch.cuda.synchronize()
    assert list(st.cpu().numpy()) == [0, 0, 1, 0, 0]
    D1 = dD1.cpu().numpy(); D2 = dD2.cpu().numpy()
    assert (D1[2] == 7.0).all() and (D2[2] == 7.0).all()      # untouched
    for f in (0, 1, 3, 4):
        R1, R2 = oracle.process(ol.robotics(dm), L[f], R[f])
        assert np.array_equal(D1[f], R1) and np.array_equal(D2[f], R2), f
    e.close()


def test_batch_is_deterministic_and_order_independent(jn, synth):
    """Size-independent property at the bench size: permuting the frames of a batch permutes
    the outputs, and two runs agree bit for bit (the segment labelling uses atomics)."""
    import torch
    W, H, dm, B = 1920, 1200, 255, 3
    L, R = synth.synth_batch(W, H, dm, [1000, 1001, 1002])
    dev = torch.device("cuda", 0)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))

    def run(order):
        dL = torch.from_numpy(L[order]).to(dev); dR = torch.from_numpy(R[order]).to(dev)
        dD = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
        e.process_batch(dL.data_ptr(), dR.data_ptr(), dD.data_ptr(), 0, 0, (W, H, W), B, 0)
        torch.cuda.synchronize()
        return dD.cpu().numpy()

    a = run([0, 1, 2]); b = run([0, 1, 2]); c = run([2, 0, 1])
    assert np.array_equal(a, b)
    assert np.array_equal(c[0], a[2]) and np.array_equal(c[1], a[0]) and np.array_equal(c[2], a[1])
    e.close()


@pytest.mark.parametrize("qname,W,H,dm,seed", [("640x480", 640, 480, 64, 1), ("1920x1200_Kx3", 1920, 1200, 255, 1000)])
def test_obstacle_scan_matches_port(jn, oracle, synth, qname, W, H, dm, seed):
    """BASELINE C2: gate cache, u8 conversion, reprojection, XR/XT, 90-bin scan, both paths.
    Gate / u8 / bin occupancy exact; ranges and angles within 1e-9 (device atan2 is 2-ulp)."""
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(fx["Q"][qname])
    A = cal.arrays()
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    D1, _ = oracle.process(ol.robotics(dm), I1, I2)
    sc = jn.ObstacleScan(cal, W, H)
    gate_ref = sp.gate(A["Q"], A["XR"], A["XT"], W, H)
    assert np.array_equal(sc.gate_cache(), gate_ref)
    u8_ref = sp.convert_u8(D1)
    r_ref, m_ref = sp.scan(A["Q"], A["XR"], A["XT"], gate_ref, u8_ref)
    r, m, u8 = sc.from_disparity(D1, want_u8=True)
    assert np.array_equal(u8, u8_ref)
    assert np.array_equal(r < 1e9 - 1, r_ref < 1e9 - 1) and m.n_finite == m_ref.n_finite
    assert m.n_points == m_ref.n_points
    assert np.allclose(r, r_ref, rtol=0, atol=1e-9)
    for k in ("angle_min", "angle_max", "range_min", "range_max"):
        assert abs(getattr(m, k) - getattr(m_ref, k)) <= 1e-9, k
    assert np.array_equal(jn.scan_compact(r), sp.compact(r))
    # -g path: full point cloud (<= 1 mm is the stated tolerance; observed exact) + scan from points
    pts_ref = sp.points(A["Q"], A["XR"], A["XT"], u8_ref)
    pts, r2, m2 = sc.points(D1)
    assert pts.shape == pts_ref.shape
    assert np.abs(pts - pts_ref).max() <= 1e-3
    r2_ref, m2_ref = sp.scan_points(pts_ref)
    assert np.array_equal(r2 < 1e9 - 1, r2_ref < 1e9 - 1)
    assert np.allclose(r2, r2_ref, rtol=0, atol=1e-9)
    # sensor_msgs/PointCloud payload: Point32 + packed rgb channel, BGR image and the grayscale quirk
    rng = np.random.default_rng(3)
    for img in (rng.integers(0, 256, D1.shape + (3,), dtype=np.uint8), rng.integers(0, 256, D1.shape, dtype=np.uint8)):
        xyz_ref, rgb_ref = sp.pointcloud(A["Q"], A["XR"], A["XT"], u8_ref, img)
        xyz, rgb, r3, m3 = sc.pointcloud(D1, img)
        assert xyz.shape == xyz_ref.shape and np.abs(xyz - xyz_ref).max() <= 1e-3
        assert np.array_equal(rgb.view(np.int32), rgb_ref.view(np.int32))
        assert np.array_equal(r3, r2)
    j, i = np.argwhere(u8_ref.T >= 2)[0][::-1]           # first point in column-major order
    assert rgb_ref.view(np.int32)[0] == (int(img[j, 3 * i + 2]) << 16 | int(img[j, 3 * i + 1]) << 8 | int(img[j, 3 * i]))
    sc.close()


def test_obstacle_scan_matches_statement_by_statement_execution(jn):
    """The device scan path against point_cloud.cpp's own statements executed with cv2 (see
    tests/golden/make_scan_statement_golden.py): gate cache incl. the 256 -> 0 wrap, u8 map, bins, point count
    exact; ranges / angles within 1e-9 (device atan2 / sqrt vs libm), -g points within 1e-9 m."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_statements.npz"))
    for name in z["names"]:
        g = lambda k: z["%s_%s" % (name, k)]
        W, H, ox, oy = [int(x) for x in g("dims")]
        cal = jn.Calibration(scan_lib.CALIB_YML)
        cal.set_q_matrix(g("Q"))
        sc = jn.ObstacleScan(cal, W, H, ox, oy)
        assert np.array_equal(sc.gate_cache(), g("gate")), name
        r, m, u8 = sc.from_disparity(g("D"), want_u8=True)
        assert np.array_equal(u8, g("u8")), name
        ref = g("scan")
        assert np.array_equal(r < 1e9 - 1, ref < 1e9 - 1), name
        assert np.allclose(r, ref, rtol=0, atol=1e-9), name
        meta = g("meta")
        assert m.n_points == int(meta[4]), name
        if m.n_points:
            for v, k in zip(meta[:4], ("angle_min", "angle_max", "range_min", "range_max")):
                assert abs(getattr(m, k) - v) <= 1e-9, (name, k)
        assert np.array_equal(jn.scan_compact(r), jn.scan_compact(ref)), name
        pts, r2, m2 = sc.points(g("D"))
        gp = g("pts")
        assert pts.shape == gp.shape, name
        fin = np.isfinite(gp).all(axis=1)
        assert np.abs(pts[fin] - gp[fin]).max() <= 1e-9, name
        rp = g("scan_p")
        assert np.array_equal(r2 < 1e9 - 1, rp < 1e9 - 1) and np.allclose(r2, rp, rtol=0, atol=1e-9), name
        sc.close()


def test_obstacle_scan_general_q_matrix(jn):
    """A Q without stereoRectify's sparsity (CALIB_ZERO_DISPARITY off: Q33 != 0, plus a skew term) takes
    the general 4x4 product; same checks against the port as the sparse one."""
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    Q = np.array(fx["Q"]["640x480"], np.float64).reshape(4, 4).copy()
    Q[3, 3] = 0.37
    Q[0, 1] = 1e-3
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(Q)
    A = cal.arrays()
    W, H = 640, 480
    rng = np.random.default_rng(5)
    D = np.floor(rng.uniform(-20, 120, (H, W))).astype(np.float32)
    sc = jn.ObstacleScan(cal, W, H)
    gate_ref = sp.gate(A["Q"], A["XR"], A["XT"], W, H)
    assert np.array_equal(sc.gate_cache(), gate_ref)
    u8_ref = sp.convert_u8(D)
    r_ref, m_ref = sp.scan(A["Q"], A["XR"], A["XT"], gate_ref, u8_ref)
    r, m, _ = sc.from_disparity(D)
    assert m.n_points == m_ref.n_points and m.n_finite == m_ref.n_finite
    assert np.array_equal(r < 1e9 - 1, r_ref < 1e9 - 1)
    assert np.allclose(r, r_ref, rtol=0, atol=1e-9)
    for k in ("angle_min", "angle_max", "range_min", "range_max"):
        assert abs(getattr(m, k) - getattr(m_ref, k)) <= 1e-9, k
    sc.close()


def test_reprojection_matches_opencv_golden(jn):
    """Device reprojection (sparse-Q path and general path) and u8 conversion against values computed by
    OpenCV itself (cv2.gemm, saturate_cast; tests/golden/scan_cv2.npz): bit for bit."""
    z = np.load(os.path.join(ol.ROOT, "tests", "golden", "scan_cv2.npz"))
    H, W = z["dmap"].shape
    cal = jn.Calibration(scan_lib.CALIB_YML)
    for qk, pk, dmap in (("Q", "pts", z["dmap"]), ("Qg", "ptsg", None)):
        if dmap is None:
            dmap = np.zeros_like(z["dmap"]); dmap[::3, ::3] = z["dmap"][::3, ::3]
        cal.set_q_matrix(z[qk])
        for i in range(9): cal.c.XR[i] = float(z["XR"].reshape(-1)[i])
        for i in range(3): cal.c.XT[i] = float(z["XT"].reshape(-1)[i])
        sc = jn.ObstacleScan(cal, W, H, int(z["ox"]), int(z["oy"]))
        pts, _, _ = sc.points(dmap.astype(np.float32))
        assert np.array_equal(pts, z[pk]), qk
        sc.close()
    Dl = z["D"]
    cal.set_q_matrix(z["Q"])
    sc = jn.ObstacleScan(cal, Dl.shape[1], 1)
    _, _, u8 = sc.from_disparity(Dl, want_u8=True)
    assert np.array_equal(u8, z["u8"])
    sc.close()


def test_scan_fast_path_is_bit_identical_to_general_expressions(jn, synth):
    """scan_kernel's fast path (table division verified at create time, float-filtered atan2) against the same
    kernel evaluating every pixel with the general double expressions: every output byte equal, on a smooth
    map, a noisy map and a map with d = 0 pixels behind a wrapped gate (H8)."""
    import ctypes as C
    fx = scan_lib.fixtures()
    rng = np.random.default_rng(5)
    for qname, W, H, oxy in (("640x480", 640, 480, (0, 0)), ("1920x1200_Kx3", 1920, 1200, (0, 0)),
                             ("640x480", 600, 400, (17, 29))):
        cal = jn.Calibration(scan_lib.CALIB_YML)
        cal.set_q_matrix(fx["Q"][qname])
        yy, xx = np.mgrid[0:H, 0:W]
        maps = [(3 + 0.02 * xx + 0.11 * yy).astype(np.float32),
                rng.uniform(-3, 260, (H, W)).astype(np.float32),
                np.where(rng.random((H, W)) < 0.3, 0, rng.integers(0, 256, (H, W))).astype(np.float32)]
        jn.lib().jn_debug_scan_fast(C.c_int(0))
        slow = jn.ObstacleScan(cal, W, H, *oxy)
        jn.lib().jn_debug_scan_fast(C.c_int(-1))
        fast = jn.ObstacleScan(cal, W, H, *oxy)
        assert slow.fast_path() == 0 and fast.fast_path() == 1
        assert np.array_equal(slow.gate_cache(), fast.gate_cache())
        for D in maps:
            r0, m0, u0 = slow.from_disparity(D, want_u8=True)
            r1, m1, u1 = fast.from_disparity(D, want_u8=True)
            assert np.array_equal(r0.view(np.int64), r1.view(np.int64))
            assert bytes(m0) == bytes(m1)
            assert np.array_equal(u0, u1)
            assert m1.n_points > 0
        slow.close(); fast.close()


def test_scan_batch_and_empty_map(jn):
    import torch
    fx = scan_lib.fixtures()
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(fx["Q"]["640x480"])
    W, H = 640, 480
    sc = jn.ObstacleScan(cal, W, H)
    r, m, _ = sc.from_disparity(np.full((H, W), -10, np.float32))   # nothing valid
    assert (r == 1e9).all() and m.n_finite == 0 and m.n_points == 0
    assert m.angle_min == 400 and m.angle_max == -400 and m.range_min == 1e9 and m.range_max == -500
    sc.close()


def test_rectify_matches_opencv_golden_and_port(jn, port):
    """jn_rectify_batch = cv::remap(INTER_LINEAR) + ROI crop, bit for bit: against the committed cv2.remap
    outputs, and against the port on a random batch with a ROI and padded strides."""
    import torch
    dev = torch.device("cuda", 0)
    for name, src, mx, my, dst in ol.remap_golden_cases():
        r = jn.Rectifier(mx, my)
        dsrc = torch.from_numpy(np.ascontiguousarray(src)).to(dev)
        dout = torch.zeros(dst.shape, dtype=torch.uint8, device=dev)
        r.remap_batch(dsrc.data_ptr(), 1, src.shape[1], src.shape[0], src.shape[1], dout.data_ptr(), dst.shape[1])
        torch.cuda.synchronize()
        assert np.array_equal(dout.cpu().numpy(), dst), name
        r.close()
    rng = np.random.default_rng(11)
    B, SH, SW, SS, H, W = 3, 70, 90, 96, 64, 80
    src = rng.integers(0, 256, (B, SH, SS), dtype=np.uint8)
    mx = rng.uniform(-2, SW + 2, (H, W)).astype(np.float32)
    my = rng.uniform(-2, SH + 2, (H, W)).astype(np.float32)
    roi = (5, 3, 61, 50)
    r = jn.Rectifier(mx, my)
    dsrc = torch.from_numpy(src).to(dev)
    dout = torch.full((B, roi[3], 64), 9, dtype=torch.uint8, device=dev)
    r.remap_batch(dsrc.data_ptr(), B, SW, SH, SS, dout.data_ptr(), 64, roi=roi)
    torch.cuda.synchronize()
    out = dout.cpu().numpy()
    for f in range(B):
        assert np.array_equal(out[f, :, :roi[2]], ol.port_remap(src[f, :, :SW], mx, my, roi)), f
    assert (out[:, :, roi[2]:] == 9).all()          # row padding untouched
    with pytest.raises(jn.JnError):
        r.remap_batch(dsrc.data_ptr(), B, SW, SH, SS, dout.data_ptr(), 64, roi=(40, 3, 61, 50))
    r.close()
