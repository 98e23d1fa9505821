"""Oracle pin of the scan side and the navigate vote against the REFERENCE'S OWN CODE: point_cloud.cpp and
navigate.cpp compiled where they lie (oracle/_ref/libpointcloud_ref.so, libnavigate_ref.so; stand-in ROS / OpenCV
headers, oracle/standins/standins.h).  The restatements the GPU tests check the CUDA path against
(oracle/scan_port.c, oracle/navigate_port.py) and the host code behind jn_navigate_* must equal what the compiled
reference functions write and publish, bit for bit, on the inputs the GPU tests use.  CPU only.

The compiled reference indexes its 90-bin array unchecked; ref_nodes_lib.safe_for_reference keeps
every input inside its defined behaviour and each test asserts that it did."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
import ref_nodes_lib as rn
import scan_lib

sys.path.insert(0, os.path.join(ol.ROOT, "oracle"))
import navigate_port  # noqa: E402

pytestmark = pytest.mark.skipif(not (os.path.exists(rn.POINTCLOUD_REF_SO) and os.path.exists(rn.NAVIGATE_REF_SO)),
                                reason="oracle/_ref node libraries not built (need /root/reference)")

FX = scan_lib.fixtures()
XR = np.array(FX["calib"]["XR"], np.float64)
XT = np.array(FX["calib"]["XT"], np.float64)


def f32(*xs):
    return [np.float32(x) for x in xs]


def meta_f32(m):
    return np.array(f32(m.angle_min, m.angle_max, m.range_min, m.range_max))


@pytest.fixture(scope="module")
def sp():
    return scan_lib.ScanPort()


@pytest.mark.parametrize("key,W,H,ox,oy", [("320x180", 320, 180, 0, 0), ("640x480", 640, 480, 0, 0),
                                           ("640x480", 200, 120, 300, 330), ("1920x1200_Kx3", 240, 150, 820, 700),
                                           ("1920x1200_Kx3", 160, 100, 0, 0)])
def test_gate_cache_equals_cacheDisparityValues(sp, key, W, H, ox, oy):
    """point_cloud.cpp:104-147, incl. crop offsets and the rows whose gate wraps to 0 (d reaches 256)."""
    Q = np.array(FX["Q"][key], np.float64)
    node = rn.PointCloudNode(Q, XR, XT, W, H, ox, oy)
    g = node.cache_gate()
    assert np.array_equal(g, sp.gate(Q, XR, XT, W, H, ox, oy))
    assert (g[..., 1] == 255).all()
    if key == "1920x1200_Kx3" and oy > 0:
        assert (g[..., 0] == 0).any() and (g[..., 0] >= 3).any()      # wrapped and ordinary pixels in one window


@pytest.mark.parametrize("W,H,key,dm,seed", [(320, 180, "320x180", 64, 3), (640, 480, "640x480", 64, 1)])
def test_image_pair_to_published_scan(sp, synth, ref, W, H, key, dm, seed):
    """imageCallbackLeft's compute on an image pair: generateDisparityMap (:406-429, the reference's Elas with its
    default parameters + postprocess_only_left, convertTo(CV_8U)) then publishPointCloud -> publishObstacleScan
    (:298-304, 213-296).  The map equals Elas::process + the port's conversion, the published LaserScan equals the
    port's scan compacted to float32."""
    Q = np.array(FX["Q"][key], np.float64)
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    node = rn.PointCloudNode(Q, XR, XT, W, H)
    gate = node.cache_gate()
    u8 = node.generate_disparity(I1, I2)
    D1, _ = ref.process(ol.robotics(255), I1, I2)
    assert np.array_equal(u8, sp.convert_u8(D1)) and (u8 > 0).mean() > 0.5
    assert rn.safe_for_reference(Q, XR, XT, u8, gate=gate)
    ranges, meta = node.scan(u8)
    pr, pm = sp.scan(Q, XR, XT, gate, u8)
    assert np.array_equal(ranges, sp.compact(pr)) and len(ranges) == pm.n_finite > 20
    assert np.array_equal(meta, meta_f32(pm))


@pytest.mark.parametrize("case", ["random", "sparse", "empty", "kx3_window"])
def test_default_scan_on_synthetic_maps(sp, case):
    """publishObstacleScan(Mat&) on maps ELAS does not produce: dense random disparities, a few isolated pixels, a
    map with nothing above the gate (empty LaserScan, angle 400 / -400, range 1e9 / -500), a window of the
    1920x1200 calibration with wrapped gate entries."""
    rng = np.random.default_rng(11)
    if case == "kx3_window":
        key, W, H, ox, oy = "1920x1200_Kx3", 240, 150, 820, 700
    else:
        key, W, H, ox, oy = "320x180", 320, 180, 0, 0
    Q = np.array(FX["Q"][key], np.float64)
    node = rn.PointCloudNode(Q, XR, XT, W, H, ox, oy)
    gate = node.cache_gate()
    if case == "random":
        u8 = rng.integers(1, 256, (H, W)).astype(np.uint8)
    elif case == "sparse":
        u8 = np.zeros((H, W), np.uint8)
        idx = rng.integers(0, W * H, 40)
        u8.flat[idx] = rng.integers(3, 200, 40)
    elif case == "empty":
        u8 = np.zeros((H, W), np.uint8)
    else:
        u8 = rng.integers(1, 256, (H, W)).astype(np.uint8)          # no 0 on a wrapped pixel: that is H8's NaN
    assert rn.safe_for_reference(Q, XR, XT, u8, gate=gate, ox=ox, oy=oy)
    ranges, meta = node.scan(u8)
    pr, pm = sp.scan(Q, XR, XT, gate, u8, ox, oy)
    assert np.array_equal(ranges, sp.compact(pr))
    assert np.array_equal(meta, meta_f32(pm))
    if case == "empty":
        assert len(ranges) == 0 and list(meta) == f32(400, -400, 1e9, -500)
    else:
        assert len(ranges) > 0


@pytest.mark.parametrize("channels", [1, 3])
def test_pointcloud_path_equals_publishPointCloud(sp, synth, ref, channels):
    """-g: publishPointCloud (:298-404) + publishObstacleScan(vector<Point3d>) (:149-211).  Point32 values, point
    order, the packed rgb channel (a grayscale frame is indexed as Vec3b, as the reference does) and the scan."""
    W, H, dm = 320, 180, 64
    Q = np.array(FX["Q"]["320x180"], np.float64)
    I1, I2, _ = synth.synth_pair(W, H, dm, 5)
    node = rn.PointCloudNode(Q, XR, XT, W, H)
    u8 = node.generate_disparity(I1, I2)
    assert rn.safe_for_reference(Q, XR, XT, u8, min_d=2)
    img = I1 if channels == 1 else np.random.default_rng(2).integers(0, 256, (H, W, 3)).astype(np.uint8)
    xyz, rgb, ranges, meta = node.pointcloud(u8, img)
    pxyz, prgb = sp.pointcloud(Q, XR, XT, u8, img)
    assert len(xyz) == len(pxyz) == int((u8 >= 2).sum()) > 1000
    assert np.array_equal(xyz, pxyz)
    assert np.array_equal(rgb.view(np.int32), prgb.view(np.int32))
    pr, pm = sp.scan_points(sp.points(Q, XR, XT, u8))
    assert np.array_equal(ranges, sp.compact(pr)) and len(ranges) > 20
    assert np.array_equal(meta, meta_f32(pm))


def _bind_nav(jn):
    l = jn.lib()
    l.jn_navigate_create.restype = C.c_void_p
    l.jn_navigate_destroy.argtypes = [C.c_void_p]
    l.jn_navigate_set_last_dir.argtypes = [C.c_void_p, C.c_int]
    l.jn_navigate_last_dir.argtypes = [C.c_void_p]
    l.jn_navigate_set_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
    l.jn_navigate_set_scan_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.jn_navigate_check_obstacle.argtypes = [C.c_void_p, C.c_void_p]
    l.jn_navigate_choose_direction.argtypes = [C.c_void_p]
    l.jn_navigate_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return l


def _same_report(fields, rep, is_obstacle):
    """checkObstacle prints `count, points, Y|N, closest, conf` with the stream's default 6 significant digits."""
    return (int(fields[0]) == rep[0] and int(fields[1]) == rep[1] and fields[2] == ("Y" if is_obstacle else "N")
            and fields[3] == "%g" % rep[2] and fields[4] == "%g" % rep[3])


def test_navigate_vote_equals_the_compiled_node(jn):
    """laserScanCallback / checkObstacle / chooseDirection (navigate.cpp:344-363, 101-153, 155-197) over 1500 scans:
    the reference's laser points equal the restatement's bit for bit, the vote, the printed report and the chosen
    direction equal jn_navigate_*'s and the restatement's, with obstacleAvoidMode's last_dir feedback."""
    l = _bind_nav(jn)
    rng = np.random.default_rng(7)
    nav = l.jn_navigate_create()
    port = navigate_port.Navigate()
    node = rn.NavigateNode()
    seen = set()
    for frame in range(1500):
        n = int(rng.integers(0, 91))
        mode = frame // 50 % 4
        if mode == 0:
            ranges = rng.uniform(0.3, 6.0, n)
        elif mode == 1:
            ranges = rng.uniform(1.2, 6.0, n)
        elif mode == 2:
            ranges = rng.uniform(0.55, 1.0, n)
        else:
            ranges = np.where(np.arange(n) < n // 2, 0.8, 5.0) * rng.uniform(0.9, 1.1, n)
        ranges = ranges.astype(np.float32)
        a0, a1 = sorted(np.float32(rng.uniform(-0.7, 0.7, 2)))          # LaserScan carries float32 angles
        node.laser_scan(ranges, a0, a1)
        assert l.jn_navigate_set_scan(nav, ranges.ctypes.data_as(C.c_void_p), n, float(a0), float(a1)) == 0
        port.laser_scan_callback(ranges, float(a0), float(a1))
        assert np.array_equal(node.points(), np.array(port.laser_points, np.float64).reshape(n, 2)), frame
        exp, fields = node.check_obstacle()
        rep = (C.c_double * 4)()
        got = l.jn_navigate_check_obstacle(nav, rep)
        pe, prep = port.check_obstacle()
        assert got == exp == pe and _same_report(fields, rep, exp) and list(rep) == list(prep), (frame, fields, list(rep))
        d = l.jn_navigate_choose_direction(nav)
        assert d == node.choose_direction() == port.choose_direction(), frame
        if frame % 3 == 0:
            l.jn_navigate_set_last_dir(nav, d); node.set_last_dir(d); port.last_dir = d
        seen.add((got, d))
    assert {0, 1} <= {s[0] for s in seen} and {0, 1, 2} <= {s[1] for s in seen}
    l.jn_navigate_destroy(nav)


def test_obstacle_avoid_mode_feedback(jn):
    """obstacleAvoidMode (navigate.cpp:229-257) is the caller that stores chooseDirection's answer in last_dir (and
    clears it when the way is free): the same loop written with jn_navigate_* keeps the same last_dir."""
    l = _bind_nav(jn)
    rng = np.random.default_rng(3)
    nav = l.jn_navigate_create()
    node = rn.NavigateNode()
    dirs = set()
    for frame in range(600):
        n = int(rng.integers(20, 91))
        side = frame // 40 % 3
        base = np.full(n, 5.0)
        if side == 1:
            base[: n // 2] = 0.8
        elif side == 2:
            base[n // 2:] = 0.8
        ranges = (base * rng.uniform(0.9, 1.1, n)).astype(np.float32)
        a0, a1 = np.float32(-0.5), np.float32(0.5)
        node.laser_scan(ranges, a0, a1)
        l.jn_navigate_set_scan(nav, ranges.ctypes.data_as(C.c_void_p), n, float(a0), float(a1))
        exp_dir, vel = node.obstacle_avoid_mode(1.0)
        if l.jn_navigate_check_obstacle(nav, None):
            l.jn_navigate_set_last_dir(nav, l.jn_navigate_choose_direction(nav))
        else:
            l.jn_navigate_set_last_dir(nav, 0)
        assert l.jn_navigate_last_dir(nav) == exp_dir, frame
        assert (vel[1] > 0) == (exp_dir == 1) and (vel[1] < 0) == (exp_dir == 2)
        dirs.add(exp_dir)
    assert dirs == {0, 1, 2}
    l.jn_navigate_destroy(nav)


def test_scan_message_into_the_vote(jn, sp, synth):
    """The two nodes chained as on the robot: point_cloud publishes a LaserScan (float32 ranges and angles), navigate
    consumes it.  jn_navigate_set_scan_bins takes the 90-bin scan + meta directly and must land on the same laser
    points: the angles go through float32 exactly as the message's fields do."""
    l = _bind_nav(jn)
    W, H, dm = 320, 180, 64
    Q = np.array(FX["Q"]["320x180"], np.float64)
    pc = rn.PointCloudNode(Q, XR, XT, W, H)
    gate = pc.cache_gate()
    node = rn.NavigateNode()
    nav = l.jn_navigate_create()
    for seed in range(4):
        I1, I2, _ = synth.synth_pair(W, H, dm, 20 + seed)
        u8 = pc.generate_disparity(I1, I2)
        assert rn.safe_for_reference(Q, XR, XT, u8, gate=gate)
        ranges, meta = pc.scan(u8)
        node.laser_scan(ranges, meta[0], meta[1])
        pr, pm = sp.scan(Q, XR, XT, gate, u8)
        m = jn.ScanMeta(pm.angle_min, pm.angle_max, pm.range_min, pm.range_max, pm.n_finite, pm.n_points)
        assert l.jn_navigate_set_scan_bins(nav, pr.ctypes.data_as(C.c_void_p), C.byref(m)) == 0
        exp, fields = node.check_obstacle()
        rep = (C.c_double * 4)()
        assert l.jn_navigate_check_obstacle(nav, rep) == exp and _same_report(fields, rep, exp)
        # the laser points themselves (what visualizeLaserPoints publishes), bit for bit
        xy = np.zeros((90, 2), np.float64)
        cnt = l.jn_navigate_points(nav, xy.ctypes.data_as(C.c_void_p), 90)
        assert cnt == len(ranges) and np.array_equal(xy[:cnt], node.points())
        port = navigate_port.Navigate()
        port.laser_scan_callback(sp.compact(pr), float(np.float32(pm.angle_min)), float(np.float32(pm.angle_max)))
        assert np.array_equal(node.points(), np.array(port.laser_points, np.float64).reshape(-1, 2))
        assert float(np.float32(pm.angle_min)) != float(pm.angle_min)    # the rounding is not a no-op on these scans
        assert l.jn_navigate_choose_direction(nav) == node.choose_direction()
    l.jn_navigate_destroy(nav)


def test_random_calibrations_and_crops(sp):
    """Forty random calibrations (focal length, principal point, baseline, a rotated and shifted camera-to-robot
    transform, a Q with every entry non-zero in a third of them) on random crops: gate cache, default scan and -g
    payload of the compiled node against the restatement.  Cases the stock node cannot take (a bin outside [1, 88]
    or a NaN angle) are counted and skipped; at least half must remain."""
    rng = np.random.default_rng(2026)
    W, H = 96, 64
    ran = 0
    for case in range(40):
        f = rng.uniform(300, 2500)
        cx, cy = rng.uniform(0.3, 0.7) * 4 * W, rng.uniform(0.3, 0.7) * 4 * H
        tx = -rng.uniform(0.05, 0.3)
        Q = np.array([[1, 0, 0, -cx], [0, 1, 0, -cy], [0, 0, 0, f], [0, 0, -1 / tx, 0]], np.float64)
        if case % 3 == 0:
            Q = Q + rng.uniform(-1e-3, 1e-3, (4, 4))                  # dense Q: every product term contributes
        a, b, c = rng.uniform(-0.08, 0.08, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        xr = Rz @ Ry @ Rx @ XR
        xt = XT.reshape(3) + rng.uniform(-0.1, 0.1, 3)
        ox, oy = int(rng.integers(0, 3 * W)), int(rng.integers(0, 3 * H))
        node = rn.PointCloudNode(Q, xr, xt, W, H, ox, oy)
        gate = node.cache_gate()
        assert np.array_equal(gate, sp.gate(Q, xr, xt, W, H, ox, oy)), case
        u8 = rng.integers(1, 256, (H, W)).astype(np.uint8)
        u8[rng.random((H, W)) < 0.3] = 0
        u8[(gate[..., 0] == 0) & (u8 == 0)] = 1                        # H8: no zero on a wrapped gate entry
        if not (rn.safe_for_reference(Q, xr, xt, u8, gate=gate, ox=ox, oy=oy)
                and rn.safe_for_reference(Q, xr, xt, u8, min_d=2, ox=ox, oy=oy)):
            continue
        ran += 1
        ranges, meta = node.scan(u8)
        pr, pm = sp.scan(Q, xr, xt, gate, u8, ox, oy)
        assert np.array_equal(ranges, sp.compact(pr)) and np.array_equal(meta, meta_f32(pm)), case
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        xyz, rgb, ranges2, meta2 = node.pointcloud(u8, img)
        pxyz, prgb = sp.pointcloud(Q, xr, xt, u8, img, ox, oy)
        assert np.array_equal(xyz, pxyz) and np.array_equal(rgb.view(np.int32), prgb.view(np.int32)), case
        pr2, pm2 = sp.scan_points(sp.points(Q, xr, xt, u8, ox, oy))
        assert np.array_equal(ranges2, sp.compact(pr2)) and np.array_equal(meta2, meta_f32(pm2)), case
    assert ran >= 20, ran


def test_python_navigate_mirror_against_the_compiled_node(jn, sp, synth):
    """jn.Navigate (the Python mirror of the navigate node's scan consumer) fed with the 90-bin scan and with the
    published LaserScan, against the compiled node fed by the compiled point_cloud node."""
    W, H, dm = 320, 180, 64
    Q = np.array(FX["Q"]["320x180"], np.float64)
    pc = rn.PointCloudNode(Q, XR, XT, W, H)
    gate = pc.cache_gate()
    node = rn.NavigateNode()
    a, b = jn.Navigate(), jn.Navigate()
    for seed in range(3):
        I1, I2, _ = synth.synth_pair(W, H, dm, 40 + seed)
        u8 = pc.generate_disparity(I1, I2)
        assert rn.safe_for_reference(Q, XR, XT, u8, gate=gate)
        ranges, meta = pc.scan(u8)
        node.laser_scan(ranges, meta[0], meta[1])
        pr, pm = sp.scan(Q, XR, XT, gate, u8)
        a.set_scan_bins(pr, jn.ScanMeta(pm.angle_min, pm.angle_max, pm.range_min, pm.range_max, pm.n_finite, pm.n_points))
        b.set_scan(ranges, meta[0], meta[1])
        assert np.array_equal(a.points(), node.points()) and np.array_equal(b.points(), node.points())
        exp, fields = node.check_obstacle()
        for nav in (a, b):
            got, rep = nav.check_obstacle()
            assert got == exp and _same_report(fields, rep, exp)
            assert nav.choose_direction() == node.choose_direction()
        d = node.choose_direction()
        node.set_last_dir(d); a.last_dir = d; b.last_dir = d
        assert a.last_dir == node.last_dir()
    with pytest.raises(ValueError):
        a.set_scan_bins(np.zeros(10), jn.ScanMeta())
    a.close(); b.close()


def test_extrinsics_from_euler_angles_equal_the_node(jn):
    """-m mode: XR = Z * Y * X from float cos / sin, XT as floats (point_cloud.cpp:76-102, 305-311).  The
    dynamic_reconfigure defaults (cfg/CamToRobotCalibParams.cfg:8-13) and 200 random settings, bit for bit."""
    node = rn.PointCloudNode(np.eye(4), np.eye(3), np.zeros(3), 8, 8)
    cal = jn.Calibration(scan_lib.CALIB_YML)
    rng = np.random.default_rng(9)
    cases = [(1.3, -3.14, 1.57, 0.0, 0.0, 0.28), (0, 0, 0, 0, 0, 0)]
    cases += [tuple(rng.uniform(-6.28, 6.28, 3)) + tuple(rng.uniform(-100, 100, 3)) for _ in range(200)]
    for c in cases:
        xr, xt = node.compose(*c)
        cal.compose_cam_to_robot(*c)
        A = cal.arrays()
        assert np.array_equal(np.asarray(A["XR"]).reshape(3, 3), xr), c
        assert np.array_equal(np.asarray(A["XT"]).reshape(3), xt), c
    assert abs(np.linalg.det(xr) - 1) < 1e-5                     # a rotation, to float precision


def test_velocity_command_equals_the_twist_the_node_publishes(jn):
    """safeNavigate (navigate.cpp:302-342) for its three working modes over 1 200 scans with changing obstacles,
    stick positions and mode switches: jn_navigate_command returns the linear.x / angular.z of the Twist the
    compiled node publishes, bit for bit (the vote, the stored direction and the acceleration ramp all carry state
    from call to call), and no button means no message."""
    rng = np.random.default_rng(12)
    node = rn.NavigateNode()
    nav = jn.Navigate()
    assert node.safe_navigate() is None                                  # no mode button: nothing published
    buttons = {jn.NAV_STOP_IN_FRONT_MANUAL: dict(r1=1, r2=1), jn.NAV_OBSTACLE_AVOID: dict(x=1),
               jn.NAV_STOP_IN_FRONT: dict(o=1)}
    seen = set()
    mode = jn.NAV_OBSTACLE_AVOID
    for frame in range(1200):
        if frame % 60 == 0:
            mode = int(rng.integers(0, 3))
        if frame == 600:
            node.set_max_forward_vel(0.9); nav.set_max_forward_vel(0.9)   # -f
        n = int(rng.integers(10, 91))
        base = np.full(n, 5.0)
        phase = frame // 45 % 4
        if phase == 1:
            base[: n // 2] = 0.8
        elif phase == 2:
            base[n // 2:] = 0.8
        elif phase == 3:
            base[:] = 0.7
        ranges = (base * rng.uniform(0.9, 1.1, n)).astype(np.float32)
        node.laser_scan(ranges, np.float32(-0.5), np.float32(0.5))
        nav.set_scan(ranges, float(np.float32(-0.5)), float(np.float32(0.5)))
        side, front = np.float32(rng.uniform(-1, 1)), np.float32(rng.uniform(-1, 1))      # Joy axes are float32
        exp = node.safe_navigate(side=side, front=front, **buttons[mode])
        got = nav.command(mode, float(side), float(front))
        assert exp is not None and got == exp, (frame, mode, got, exp)
        assert nav.last_dir == node.last_dir()
        seen.add((mode, got[0] > 0, got[1] > 0, got[1] < 0))
    assert len(seen) >= 8                                                # forward, stopped, turning either way
    with pytest.raises(jn.JnError):
        nav.command(7)
    nav.close()


@pytest.mark.parametrize("qname,W,H,dm,seed,rows", [("640x480", 640, 480, 64, 1, 480), ("1920x1200_Kx3", 1920, 1200, 255, 1000, 320)])
def test_inputs_of_the_gpu_scan_test_through_the_compiled_node(sp, synth, ref, qname, W, H, dm, seed, rows):
    """The maps tests/test_gpu_parity.py::test_obstacle_scan_matches_port hands the CUDA path (same scene, seed,
    parameters and calibration), handed to the compiled node instead: the restatement that test compares against equals
    the node on them, default path and -g path.  At 1920x1200 the stock node cannot take the whole map (the gate of the
    rows below ~775 wraps to 0 and an invalid pixel there is H8's NaN angle, asserted here); it gets the rows above."""
    Q = np.array(FX["Q"][qname], np.float64)
    I1, I2, _ = synth.synth_pair(W, H, dm, seed)
    D1, _ = ref.process(ol.robotics(dm), I1, I2)
    u8_full = sp.convert_u8(D1)
    if rows < H:
        g_full = sp.gate(Q, XR, XT, W, H)
        assert not rn.safe_for_reference(Q, XR, XT, u8_full, gate=g_full)
        assert rn.bin_index_range(Q, XR, XT, u8_full, gate=g_full)[3] > 0          # NaN angles: the reason
    u8 = np.ascontiguousarray(u8_full[:rows])
    node = rn.PointCloudNode(Q, XR, XT, W, rows)
    gate = node.cache_gate()
    assert np.array_equal(gate, sp.gate(Q, XR, XT, W, rows))
    assert rn.safe_for_reference(Q, XR, XT, u8, gate=gate) and rn.safe_for_reference(Q, XR, XT, u8, min_d=2)
    ranges, meta = node.scan(u8)
    pr, pm = sp.scan(Q, XR, XT, gate, u8)
    assert np.array_equal(ranges, sp.compact(pr)) and np.array_equal(meta, meta_f32(pm)) and len(ranges) > 20
    xyz, rgb, ranges2, meta2 = node.pointcloud(u8, np.ascontiguousarray(I1[:rows]))
    pxyz, prgb = sp.pointcloud(Q, XR, XT, u8, np.ascontiguousarray(I1[:rows]))
    assert np.array_equal(xyz, pxyz) and np.array_equal(rgb.view(np.int32), prgb.view(np.int32))
    pr2, pm2 = sp.scan_points(sp.points(Q, XR, XT, u8))
    assert np.array_equal(ranges2, sp.compact(pr2)) and np.array_equal(meta2, meta_f32(pm2))
