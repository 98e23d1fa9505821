"""ctypes access to oracle/scan_port.c (TEST INFRASTRUCTURE ONLY) + calibration fixtures."""
import ctypes as C
import json
import os
import numpy as np
from oracle_lib import PORT_SO, ROOT

FIX = os.path.join(ROOT, "tests", "golden", "q_fixtures.json")
CALIB_YML = os.path.join(ROOT, "tests", "golden", "calib_c920.yml")


class Meta(C.Structure):
    _fields_ = [("angle_min", C.c_double), ("angle_max", C.c_double), ("range_min", C.c_double),
                ("range_max", C.c_double), ("n_finite", C.c_int32), ("n_points", C.c_int32)]


def fixtures():
    return json.load(open(FIX))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ScanPort:
    def __init__(self):
        self.lib = C.CDLL(PORT_SO)

    def gate(self, Q, XR, XT, W, H, ox=0, oy=0):
        Q = np.ascontiguousarray(Q, np.float64); XR = np.ascontiguousarray(XR, np.float64)
        XT = np.ascontiguousarray(XT, np.float64)
        g = np.zeros((H, W, 2), np.uint8)
        self.lib.port_gate_cache(_p(Q), _p(XR), _p(XT), W, H, ox, oy, _p(g))
        return g

    def convert_u8(self, D):
        D = np.ascontiguousarray(D, np.float32)
        out = np.zeros(D.shape, np.uint8)
        self.lib.port_convert_u8(_p(D), C.c_size_t(D.size), _p(out))
        return out

    def scan(self, Q, XR, XT, gate, dmap, ox=0, oy=0):
        H, W = dmap.shape
        Q = np.ascontiguousarray(Q, np.float64); XR = np.ascontiguousarray(XR, np.float64)
        XT = np.ascontiguousarray(XT, np.float64)
        dmap = np.ascontiguousarray(dmap, np.uint8); gate = np.ascontiguousarray(gate, np.uint8)
        ranges = np.zeros(90, np.float64)
        m = Meta()
        self.lib.port_scan_from_dmap(_p(Q), _p(XR), _p(XT), _p(gate), _p(dmap), W, H, ox, oy, _p(ranges), C.byref(m))
        return ranges, m

    def points(self, Q, XR, XT, dmap, ox=0, oy=0):
        H, W = dmap.shape
        Q = np.ascontiguousarray(Q, np.float64); XR = np.ascontiguousarray(XR, np.float64)
        XT = np.ascontiguousarray(XT, np.float64)
        dmap = np.ascontiguousarray(dmap, np.uint8)
        pts = np.zeros((W * H, 3), np.float64)
        f = self.lib.port_points_from_dmap
        f.restype = C.c_int
        n = f(_p(Q), _p(XR), _p(XT), _p(dmap), W, H, ox, oy, _p(pts))
        return pts[:n]

    def pointcloud(self, Q, XR, XT, dmap, image, ox=0, oy=0):
        """(xyz float32 n x 3, rgb float32 n) as publishPointCloud fills sensor_msgs/PointCloud."""
        H, W = dmap.shape
        Q = np.ascontiguousarray(Q, np.float64); XR = np.ascontiguousarray(XR, np.float64)
        XT = np.ascontiguousarray(XT, np.float64)
        dmap = np.ascontiguousarray(dmap, np.uint8); image = np.ascontiguousarray(image, np.uint8)
        xyz = np.zeros((W * H, 3), np.float32); rgb = np.zeros(W * H, np.float32)
        f = self.lib.port_pointcloud_pack
        f.restype = C.c_int
        n = f(_p(Q), _p(XR), _p(XT), _p(dmap), W, H, ox, oy, _p(image), image.strides[0], 3 if image.ndim == 3 else 1,
              _p(xyz), _p(rgb))
        return xyz[:n], rgb[:n]

    def scan_points(self, pts):
        pts = np.ascontiguousarray(pts, np.float64)
        ranges = np.zeros(90, np.float64)
        m = Meta()
        self.lib.port_scan_from_points(_p(pts), pts.shape[0], _p(ranges), C.byref(m))
        return ranges, m

    def compact(self, ranges):
        ranges = np.ascontiguousarray(ranges, np.float64)
        out = np.zeros(90, np.float32)
        f = self.lib.port_scan_compact
        f.restype = C.c_int
        n = f(_p(ranges), _p(out))
        return out[:n]
