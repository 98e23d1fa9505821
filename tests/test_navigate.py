"""jn_navigate_* (host code of the C ABI, no GPU needed) against the plain-Python restatement of
navigate.cpp's laserScanCallback / checkObstacle / chooseDirection (oracle/navigate_port.py)."""
import ctypes as C
import os
import sys
import numpy as np
import oracle_lib as ol

sys.path.insert(0, os.path.join(ol.ROOT, "oracle"))
import navigate_port


def bind(jn):
    l = jn.lib()
    l.jn_navigate_create.restype = C.c_void_p
    l.jn_navigate_destroy.argtypes = [C.c_void_p]
    l.jn_navigate_set_last_dir.argtypes = [C.c_void_p, C.c_int]
    l.jn_navigate_set_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
    l.jn_navigate_set_scan_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.jn_navigate_check_obstacle.argtypes = [C.c_void_p, C.c_void_p]
    l.jn_navigate_choose_direction.argtypes = [C.c_void_p]
    return l


def test_vote_and_direction_match_the_restatement(jn):
    l = bind(jn)
    rng = np.random.default_rng(5)
    nav = l.jn_navigate_create()
    ref = navigate_port.Navigate()
    seen = set()
    for frame in range(400):
        n = int(rng.integers(0, 91))
        mode = frame // 50 % 4
        if mode == 0:
            ranges = rng.uniform(0.3, 6.0, n)
        elif mode == 1:
            ranges = rng.uniform(1.2, 6.0, n)                 # nothing inside the box: the vote decays
        elif mode == 2:
            ranges = rng.uniform(0.55, 1.0, n)                # a wall in front
        else:
            ranges = np.where(np.arange(n) < n // 2, 0.8, 5.0) * rng.uniform(0.9, 1.1, n)   # one-sided obstacle
        ranges = ranges.astype(np.float32)
        a0, a1 = sorted(rng.uniform(-0.7, 0.7, 2))
        assert l.jn_navigate_set_scan(nav, ranges.ctypes.data_as(C.c_void_p), n, a0, a1) == 0
        ref.laser_scan_callback(ranges, a0, a1)
        rep = (C.c_double * 4)()
        got = l.jn_navigate_check_obstacle(nav, rep)
        exp, (count, npts, closest, conf) = ref.check_obstacle()
        assert got == exp and list(rep) == [count, npts, closest, conf], frame
        d = l.jn_navigate_choose_direction(nav)
        assert d == ref.choose_direction(), frame
        if frame % 3 == 0:                                    # obstacleAvoidMode stores its choice
            l.jn_navigate_set_last_dir(nav, d); ref.last_dir = d
        seen.add((got, d))
    assert {0, 1} <= {s[0] for s in seen} and {0, 1, 2} <= {s[1] for s in seen}
    l.jn_navigate_destroy(nav)


def test_scan_bins_feed_the_vote(jn):
    """The 90-bin scan goes in as the reference publishes it: finite bins, k = 89..0, float32."""
    l = bind(jn)
    ranges = np.full(90, 1e9)
    ranges[[10, 11, 12, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49]] = np.linspace(0.6, 0.9, 13)
    meta = jn.ScanMeta(-0.4, 0.5, 0.6, 0.9, 13, 100)
    nav = l.jn_navigate_create()
    assert l.jn_navigate_set_scan_bins(nav, ranges.ctypes.data_as(C.c_void_p), C.byref(meta)) == 0
    ref = navigate_port.Navigate()
    ref.laser_scan_callback(jn.scan_compact(ranges), -0.4, 0.5)
    rep = (C.c_double * 4)()
    assert l.jn_navigate_check_obstacle(nav, rep) == ref.check_obstacle()[0] == 1
    assert rep[1] == 13
    assert l.jn_navigate_choose_direction(nav) == ref.choose_direction()
    l.jn_navigate_destroy(nav)
