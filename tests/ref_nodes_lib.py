"""ctypes access to oracle/_ref/libpointcloud_ref.so and libnavigate_ref.so (TEST INFRASTRUCTURE ONLY): the
reference's point_cloud.cpp and navigate.cpp compiled where they lie against oracle/standins, see
oracle/pointcloud_ref_shim.cpp / navigate_ref_shim.cpp."""
import ctypes as C
import os

import numpy as np

from oracle_lib import ROOT

POINTCLOUD_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpointcloud_ref.so")
NAVIGATE_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnavigate_ref.so")
P = C.c_void_p


def _p(a):
    return a.ctypes.data_as(P)


class PointCloudNode:
    """The node's file-scope state after main()'s set-up (point_cloud.cpp:546-559) for one calibration and crop."""

    def __init__(self, Q, XR, XT, W, H, ox=0, oy=0):
        l = C.CDLL(POINTCLOUD_REF_SO)
        l.ref_pc_setup.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ref_pc_cache_gate.argtypes = [P]
        l.ref_pc_set_gate.argtypes = [P]
        l.ref_pc_scan.argtypes = [P, P, P, P]
        l.ref_pc_pointcloud.argtypes = [P, P, C.c_int, C.c_int, P, P, P, P, P, P]
        l.ref_pc_generate_disparity.argtypes = [P, P, P]
        self.l, self.W, self.H = l, W, H
        self.Q = np.ascontiguousarray(Q, np.float64)
        self.XR = np.ascontiguousarray(XR, np.float64)
        self.XT = np.ascontiguousarray(XT, np.float64)
        l.ref_pc_setup(_p(self.Q), _p(self.XR), _p(self.XT), W, H, ox, oy)

    def compose(self, phi_x, phi_y, phi_z, tx, ty, tz):
        """composeRotationCamToRobot / composeTranslationCamToRobot -> (XR 3x3, XT 3)."""
        self.l.ref_pc_compose.argtypes = [C.c_double] * 6 + [P, P]
        xr, xt = np.zeros(9, np.float64), np.zeros(3, np.float64)
        self.l.ref_pc_compose(phi_x, phi_y, phi_z, tx, ty, tz, _p(xr), _p(xt))
        return xr.reshape(3, 3), xt

    def cache_gate(self):
        """cacheDisparityValues() -> valid_disp as H x W x 2 bytes."""
        g = np.zeros((self.H, self.W, 2), np.uint8)
        self.l.ref_pc_cache_gate(_p(g))
        return g

    def set_gate(self, gate):
        gate = np.ascontiguousarray(gate, np.uint8)
        assert gate.shape == (self.H, self.W, 2)
        self.l.ref_pc_set_gate(_p(gate))

    def scan(self, dmap_u8):
        """publishPointCloud(dmap) without -g -> (LaserScan.ranges float32, (angle_min, angle_max, range_min,
        range_max) float32)."""
        dmap_u8 = np.ascontiguousarray(dmap_u8, np.uint8)
        assert dmap_u8.shape == (self.H, self.W)
        ranges = np.zeros(90, np.float32)
        n = C.c_int32()
        meta = np.zeros(4, np.float32)
        self.l.ref_pc_scan(_p(dmap_u8), _p(ranges), C.byref(n), _p(meta))
        return ranges[:n.value].copy(), meta

    def pointcloud(self, dmap_u8, image):
        """publishPointCloud(dmap) with -g -> (Point32 xyz n x 3, rgb channel n, ranges, meta)."""
        dmap_u8 = np.ascontiguousarray(dmap_u8, np.uint8)
        image = np.ascontiguousarray(image, np.uint8)
        xyz = np.zeros((self.W * self.H, 3), np.float32)
        rgb = np.zeros(self.W * self.H, np.float32)
        ranges = np.zeros(90, np.float32)
        npts, n = C.c_int32(), C.c_int32()
        meta = np.zeros(4, np.float32)
        self.l.ref_pc_pointcloud(_p(dmap_u8), _p(image), image.strides[0], 3 if image.ndim == 3 else 1, _p(xyz), _p(rgb),
                                 C.byref(npts), _p(ranges), C.byref(n), _p(meta))
        return xyz[:npts.value].copy(), rgb[:npts.value].copy(), ranges[:n.value].copy(), meta

    def generate_disparity(self, I1, I2):
        """generateDisparityMap(left, right) -> the CV_8U map the node publishes and scans."""
        I1 = np.ascontiguousarray(I1, np.uint8)
        I2 = np.ascontiguousarray(I2, np.uint8)
        assert I1.shape == I2.shape == (self.H, self.W)
        out = np.zeros((self.H, self.W), np.uint8)
        self.l.ref_pc_generate_disparity(_p(I1), _p(I2), _p(out))
        return out


class NavigateNode:
    def __init__(self):
        l = C.CDLL(NAVIGATE_REF_SO)
        l.ref_nav_laser_scan.argtypes = [P, C.c_int, C.c_float, C.c_float]
        l.ref_nav_check_obstacle.argtypes = [C.c_char_p, C.c_int]
        l.ref_nav_points.argtypes = [P]
        l.ref_nav_set_clearance.argtypes = [C.c_double, C.c_double, C.c_int]
        l.ref_nav_obstacle_avoid_mode.argtypes = [C.c_double, P]
        self.l = l
        l.ref_nav_reset()

    def laser_scan(self, ranges_f32, angle_min, angle_max):
        r = np.ascontiguousarray(ranges_f32, np.float32)
        self.l.ref_nav_laser_scan(_p(r), len(r), float(angle_min), float(angle_max))

    def points(self):
        n = self.l.ref_nav_point_count()
        xy = np.zeros((n, 2), np.float64)
        self.l.ref_nav_points(_p(xy))
        return xy

    def check_obstacle(self):
        """(isObstacle, the fields of the line checkObstacle prints: count, points, 'Y'|'N', closest, conf)"""
        buf = C.create_string_buffer(256)
        r = self.l.ref_nav_check_obstacle(buf, 256)
        return r, buf.value.decode().strip().split(", ")

    def choose_direction(self):
        return self.l.ref_nav_choose_direction()

    def set_last_dir(self, d):
        self.l.ref_nav_set_last_dir(int(d))

    def last_dir(self):
        return self.l.ref_nav_last_dir()

    def safe_navigate(self, r1=0, r2=0, x=0, o=0, side=0.0, front=0.0):
        """safeNavigate on a Joy message -> (linear.x, angular.z) of the published Twist, or None."""
        self.l.ref_nav_safe_navigate.argtypes = [C.c_int] * 4 + [C.c_float, C.c_float, P]
        v = np.zeros(2, np.float64)
        return tuple(v) if self.l.ref_nav_safe_navigate(r1, r2, x, o, side, front, _p(v)) else None

    def set_max_forward_vel(self, v):
        self.l.ref_nav_set_max_forward_vel.argtypes = [C.c_float]
        self.l.ref_nav_set_max_forward_vel(v)

    def obstacle_avoid_mode(self, front):
        v = np.zeros(2, np.float64)
        d = self.l.ref_nav_obstacle_avoid_mode(float(front), _p(v))
        return d, v


def bin_index_range(Q, XR, XT, dmap, gate=None, min_d=None, ox=0, oy=0):
    """Safety check before handing a map to the compiled reference, which indexes scan[k] unchecked:
    (pixels selected, min k, max k, NaN angles) over the pixels the default path (gate) or the -g path (min_d)
    keeps.  Plain numpy; only used to keep the inputs inside the reference's defined behaviour."""
    H, W = dmap.shape
    jj, ii = np.mgrid[0:H, 0:W]
    d = dmap.astype(np.float64)
    sel = (dmap >= gate[..., 0]) & (dmap <= gate[..., 1]) if min_d is None else dmap >= min_d
    V = np.stack([ii + ox, jj + oy, d, np.ones_like(d)], -1)[sel]
    if len(V) == 0:
        return 0, None, None, 0
    with np.errstate(all="ignore"):
        pos = V @ np.asarray(Q, np.float64).reshape(4, 4).T
        pt = pos[:, :3] / pos[:, 3:4]
        rb = pt @ np.asarray(XR, np.float64).reshape(3, 3).T + np.asarray(XT, np.float64).reshape(1, 3)
        k = np.floor(90 * (45 - np.arctan2(rb[:, 1], rb[:, 0]) * 180 / 3.1415) / 90)
    nan = int(np.isnan(k).sum())
    return int(sel.sum()), (None if nan == len(k) else float(np.nanmin(k))), (None if nan == len(k) else float(np.nanmax(k))), nan


def safe_for_reference(Q, XR, XT, dmap, gate=None, min_d=None, ox=0, oy=0):
    n, lo, hi, nan = bin_index_range(Q, XR, XT, dmap, gate, min_d, ox, oy)
    return n == 0 or (nan == 0 and lo >= 1 and hi <= 88)
