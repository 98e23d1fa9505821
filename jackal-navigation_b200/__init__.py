"""jackal-navigation_b200 -- B200-native stereo-to-obstacle hot path.

Host-side mirror of the reference interface for this path:

    Elas.parameters / Elas(param).process(I1, I2, D1, D2, dims)   src/elas/elas.h:59-162
    calibration YAML (K1,K2,D1,D2,R,T,XR,XT) + Q                  point_cloud.cpp:530-544
    generateDisparityMap / publishObstacleScan                    point_cloud.cpp:406-429, 213-296

All compute happens in libjn_elas.so (hand-written CUDA for sm_100a) behind the C ABI
declared in include/jn_elas.h; this module only marshals pointers with ctypes.  There is
no CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
PyTorch is used by callers for device memory and streams only.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JN_ELAS_LIB") or os.path.join(_HERE, "libjn_elas.so")   # override: kernel tuning sweeps

ROBOTICS, MIDDLEBURY = 0, 1
JN_OK, JN_FEW_SUPPORT = 0, 1
SCAN_BINS = 90
SCAN_INF = 1e9


class JnError(RuntimeError):
    pass


class parameters(C.Structure):
    """Elas::parameters (elas.h:59-145): same 23 fields, same order."""
    _fields_ = [
        ("disp_min", C.c_int32), ("disp_max", C.c_int32),
        ("support_threshold", C.c_float), ("support_texture", C.c_int32),
        ("candidate_stepsize", C.c_int32), ("incon_window_size", C.c_int32),
        ("incon_threshold", C.c_int32), ("incon_min_support", C.c_int32),
        ("add_corners", C.c_int32), ("grid_size", C.c_int32),
        ("beta", C.c_float), ("gamma", C.c_float), ("sigma", C.c_float),
        ("sradius", C.c_float), ("match_texture", C.c_int32),
        ("lr_threshold", C.c_int32), ("speckle_sim_threshold", C.c_float),
        ("speckle_size", C.c_int32), ("ipol_gap_width", C.c_int32),
        ("filter_median", C.c_int32), ("filter_adaptive_mean", C.c_int32),
        ("postprocess_only_left", C.c_int32), ("subsampling", C.c_int32),
    ]

    def __init__(self, setting=ROBOTICS, **overrides):
        super().__init__()
        lib().jn_elas_params_default(C.byref(self), int(setting))
        for k, v in overrides.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class Calib(C.Structure):
    _fields_ = [("K1", C.c_double * 9), ("K2", C.c_double * 9), ("D1", C.c_double * 5), ("D2", C.c_double * 5),
                ("R", C.c_double * 9), ("T", C.c_double * 3), ("XR", C.c_double * 9), ("XT", C.c_double * 3),
                ("Q", C.c_double * 16), ("has_q", C.c_int32)]


class ScanMeta(C.Structure):
    _fields_ = [("angle_min", C.c_double), ("angle_max", C.c_double), ("range_min", C.c_double),
                ("range_max", C.c_double), ("n_finite", C.c_int32), ("n_points", C.c_int32)]


_P = C.c_void_p


class StageDump(C.Structure):
    """include/jn_elas_debug.h: jn_stage_dump."""
    _fields_ = [
        ("desc1", _P), ("desc2", _P),
        ("dcan_raw", _P), ("dcan_incon", _P), ("dcan_final", _P),
        ("support", _P), ("cap_support", C.c_int32), ("n_support", C.c_int32),
        ("tri1", _P), ("planes1", _P), ("tri2", _P), ("planes2", _P),
        ("cap_tri", C.c_int32), ("n_tri1", C.c_int32), ("n_tri2", C.c_int32),
        ("grid1", _P), ("grid2", _P),
        ("D1_raw", _P), ("D2_raw", _P), ("D1_lr", _P), ("D2_lr", _P),
        ("D1_seg", _P), ("D2_seg", _P), ("D1_gap", _P), ("D2_gap", _P),
        ("D1_mean", _P), ("D2_mean", _P), ("D1", _P), ("D2", _P),
        ("dense_evals", C.c_int64), ("dense_pixels", C.c_int64),
    ]


_lib = None


def lib():
    """Loads libjn_elas.so.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JnError("%s not built: run `python jackal-navigation_b200/build.py` "
                          "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.jn_last_error.restype = C.c_char_p
        l.jn_launch_count.restype = C.c_longlong
        l.jn_elas_create.restype = _P
        l.jn_elas_create.argtypes = [C.POINTER(parameters), C.c_int]
        l.jn_elas_destroy.argtypes = [_P]
        l.jn_elas_process.argtypes = [_P, _P, _P, _P, _P, C.POINTER(C.c_int32)]
        l.jn_elas_process_batch.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P, C.POINTER(C.c_int32), _P]
        l.jn_elas_stages.argtypes = [_P, _P, _P, C.POINTER(C.c_int32), C.POINTER(StageDump)]
        l.jn_calib_load_yaml.argtypes = [C.c_char_p, C.POINTER(Calib)]
        l.jn_calib_set_q.argtypes = [C.POINTER(Calib), C.c_double, C.c_double, C.c_double, C.c_double]
        l.jn_scan_create.restype = _P
        l.jn_scan_create.argtypes = [C.POINTER(Calib), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        l.jn_scan_destroy.argtypes = [_P]
        l.jn_scan_gate_cache.argtypes = [_P, _P]
        l.jn_scan_from_disparity_batch.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P]
        l.jn_scan_from_disparity.argtypes = [_P, _P, _P, C.POINTER(ScanMeta), _P]
        l.jn_points_from_disparity.argtypes = [_P, _P, _P, C.POINTER(C.c_int32), _P, C.POINTER(ScanMeta)]
        l.jn_scan_compact.argtypes = [_P, _P]
        l.jn_pointcloud_from_disparity.argtypes = [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.POINTER(C.c_int32), _P,
                                                   C.POINTER(ScanMeta)]
        l.jn_cache_clear.restype = None
        l.jn_host_alloc.restype = _P
        l.jn_host_alloc.argtypes = [C.c_size_t]
        l.jn_host_free.argtypes = [_P]
        _ss = [_P, _P, C.c_int, _P, _P, C.POINTER(C.c_int32), _P, _P, _P, _P, _P]
        l.jn_stereo_scan_batch_host.argtypes = _ss
        l.jn_stereo_scan_submit.argtypes = _ss
        l.jn_stereo_scan_submit_device.argtypes = _ss
        l.jn_stereo_scan_wait.argtypes = [_P]
        l.jn_rectify_create.restype = _P
        l.jn_rectify_create.argtypes = [_P, _P, C.c_int32, C.c_int32, C.c_int32]
        l.jn_rectify_destroy.argtypes = [_P]
        l.jn_rectify_batch.argtypes = [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P]
        _lib = l
    return _lib


def last_error():
    return lib().jn_last_error().decode()


def launch_count():
    return int(lib().jn_launch_count())


def _check(rc, what):
    if rc < 0:
        raise JnError("%s failed (%d): %s" % (what, rc, last_error()))
    return rc


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def _need(a, count, what):
    """The C side reads `count` elements through a raw pointer: an undersized array is a host out-of-bounds read."""
    if a.size < count:
        raise ValueError("%s must hold at least %d elements, got %d" % (what, count, a.size))
    return a


class Elas:
    """Mirror of class Elas (elas.h:52-234)."""
    ROBOTICS, MIDDLEBURY = ROBOTICS, MIDDLEBURY
    parameters = parameters

    def __init__(self, param=None, device=0):
        self.param = param if param is not None else parameters()
        self._h = lib().jn_elas_create(C.byref(self.param), int(device))
        if not self._h:
            raise JnError("jn_elas_create: " + last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().jn_elas_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, I1, I2, D1, D2, dims):
        """Elas::process(I1, I2, D1, D2, dims) with numpy host arrays (elas.h:162).

        Returns JN_OK or JN_FEW_SUPPORT; in the latter case D1/D2 are untouched and the
        reference's message is printed (elas.cpp:66-71).  With param.subsampling the maps are
        (H/2) x (W/2) (elas.h:160-162)."""
        for a, dt in ((I1, np.uint8), (I2, np.uint8), (D1, np.float32)) + (((D2, np.float32),) if D2 is not None else ()):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                raise TypeError("process expects C-contiguous numpy arrays (uint8 images, float32 maps)")
        d = (C.c_int32 * 3)(*[int(x) for x in dims])
        if d[2] < d[0] or d[0] <= 0 or d[1] <= 0:
            raise ValueError("dims = (width, height, bytes_per_line) with bytes_per_line >= width")
        if I1.size < d[2] * d[1] or I2.size < d[2] * d[1]:
            raise ValueError("images must hold height * bytes_per_line = %d bytes" % (d[2] * d[1]))
        Hd, Wd = self.map_shape(d[0], d[1])
        if D1.size < Hd * Wd or (D2 is not None and D2.size < Hd * Wd):
            raise ValueError("disparity maps must hold %d x %d floats" % (Hd, Wd))
        rc = _check(lib().jn_elas_process(self._h, _ptr(I1), _ptr(I2), _ptr(D1), _ptr(D2), d), "jn_elas_process")
        if rc == JN_FEW_SUPPORT:
            print("ERROR: Need at least 3 support points!")
        return rc

    def stereo_scan_submit(self, scan, n, I1, I2, dims, ranges, meta, status=0, dmap_u8=0, D1=0, wait=False):
        """n frames from HOST buffers (addresses as ints, ideally pinned: jn_host_alloc or
        torch.Tensor.pin_memory) through ELAS + obstacle scan; results land in host buffers.
        Asynchronous unless wait=True; see jn_stereo_scan_submit in include/jn_elas.h."""
        d = (C.c_int32 * 3)(*[int(x) for x in dims])
        f = lib().jn_stereo_scan_batch_host if wait else lib().jn_stereo_scan_submit
        P = lambda x: _P(int(x)) if x else None
        return _check(f(self._h, scan._h, int(n), P(I1), P(I2), d, P(D1), P(status), P(ranges), P(meta), P(dmap_u8)),
                      "jn_stereo_scan_submit")

    def stereo_scan_submit_device(self, scan, n, I1, I2, dims, D1, status, ranges, meta, dmap_u8=0):
        """Like stereo_scan_submit with every buffer in DEVICE memory (addresses as ints); no copies, no
        caller stream: inputs complete at call time, results valid after stereo_scan_wait()."""
        d = (C.c_int32 * 3)(*[int(x) for x in dims])
        P = lambda x: _P(int(x)) if x else None
        return _check(lib().jn_stereo_scan_submit_device(self._h, scan._h, int(n), P(I1), P(I2), d, P(D1), P(status),
                                                         P(ranges), P(meta), P(dmap_u8)),
                      "jn_stereo_scan_submit_device")

    def stereo_scan_wait(self):
        return _check(lib().jn_stereo_scan_wait(self._h), "jn_stereo_scan_wait")

    def map_shape(self, W, H):
        """(rows, cols) of the disparity maps for W x H images."""
        return (H // 2, W // 2) if self.param.subsampling else (H, W)

    def process_batch(self, I1, I2, D1, D2, status, dims, n, stream=0):
        """Device pointers (ints), frame-major batch of n frames; asynchronous on `stream`."""
        d = (C.c_int32 * 3)(*[int(x) for x in dims])
        return _check(lib().jn_elas_process_batch(self._h, int(n), _P(I1), _P(I2), _P(D1), _P(D2) if D2 else None,
                                                  _P(status) if status else None, d, _P(stream) if stream else None),
                      "jn_elas_process_batch")

    def stages(self, I1, I2, want_desc=True, want_grid=True):
        """Runs the device pipeline on one frame and returns every intermediate (tests only)."""
        H, W = I1.shape
        p = self.param
        step, gs = p.candidate_stepsize, p.grid_size
        if p.subsampling:
            step += step % 2          # elas.cpp:379-381
        Hd, Wd = self.map_shape(W, H)
        Wc, Hc = (W + step - 1) // step, (H + step - 1) // step
        gw = int(np.ceil(np.float32(W) / np.float32(gs)))
        gh = int(np.ceil(np.float32(H) / np.float32(gs)))
        I1 = np.ascontiguousarray(I1)
        I2 = np.ascontiguousarray(I2)
        cap_s, cap_t = Wc * Hc + 8, 2 * (Wc * Hc + 8)
        o = {}
        if want_desc:
            o["desc1"] = np.zeros((H, W, 16), np.uint8)
            o["desc2"] = np.zeros((H, W, 16), np.uint8)
        for k in ("dcan_raw", "dcan_incon", "dcan_final"):
            o[k] = np.zeros((Hc, Wc), np.int16)
        o["support"] = np.zeros((cap_s, 3), np.int32)
        o["tri1"] = np.zeros((cap_t, 3), np.int32)
        o["tri2"] = np.zeros((cap_t, 3), np.int32)
        o["planes1"] = np.zeros((cap_t, 6), np.float32)
        o["planes2"] = np.zeros((cap_t, 6), np.float32)
        if want_grid:
            o["grid1"] = np.zeros((gh, gw, p.disp_max + 2), np.int32)
            o["grid2"] = np.zeros((gh, gw, p.disp_max + 2), np.int32)
        for k in ("D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap",
                  "D1_mean", "D2_mean", "D1", "D2"):
            o[k] = np.zeros((Hd, Wd), np.float32)
        st = StageDump()
        for k, a in o.items():
            setattr(st, k, _ptr(a))
        st.cap_support, st.cap_tri = cap_s, cap_t
        d = (C.c_int32 * 3)(W, H, W)
        rc = _check(lib().jn_elas_stages(self._h, _ptr(I1), _ptr(I2), d, C.byref(st)), "jn_elas_stages")
        o["rc"] = rc
        o["n_support"] = st.n_support
        o["support"] = o["support"][:st.n_support]
        if rc == 0:
            o["tri1"], o["tri2"] = o["tri1"][:st.n_tri1], o["tri2"][:st.n_tri2]
            o["planes1"], o["planes2"] = o["planes1"][:st.n_tri1], o["planes2"][:st.n_tri2]
        return o


def debug_support_filter(elas, dcan, W, H):
    """(tests) support filtering + compaction on an injected candidate image."""
    dcan = np.ascontiguousarray(dcan, np.int16)
    Hc, Wc = dcan.shape
    inc = np.zeros_like(dcan); fin = np.zeros_like(dcan)
    sup = np.zeros((Hc * Wc + 8, 3), np.int32)
    n = C.c_int32(0); rounds = C.c_int32(0)
    d = (C.c_int32 * 3)(W, H, W)
    f = lib().jn_debug_support_filter
    f.argtypes = [_P, _P, C.POINTER(C.c_int32), _P, _P, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    _check(f(elas._h, _ptr(dcan), d, _ptr(inc), _ptr(fin), _ptr(sup), sup.shape[0], C.byref(n), C.byref(rounds)),
           "jn_debug_support_filter")
    return inc, fin, sup[:n.value].copy(), rounds.value


def debug_triangulate(elas, xy, W=640, H=480):
    """(tests) the Delaunay kernel on injected integer points."""
    xy = np.ascontiguousarray(xy, np.int32)
    n = xy.shape[0]
    tri = np.zeros((2 * n + 8, 3), np.int32)
    nt = C.c_int32(0)
    d = (C.c_int32 * 3)(W, H, W)
    f = lib().jn_debug_triangulate
    f.argtypes = [_P, _P, C.c_int, C.POINTER(C.c_int32), _P, C.c_int32, C.POINTER(C.c_int32)]
    _check(f(elas._h, _ptr(xy), n, d, _ptr(tri), tri.shape[0], C.byref(nt)), "jn_debug_triangulate")
    return tri[:nt.value].copy()


def debug_postprocess(elas, D1raw, D2raw):
    """(tests) the post-processing chain on injected raw disparity maps."""
    D1raw = np.ascontiguousarray(D1raw, np.float32); D2raw = np.ascontiguousarray(D2raw, np.float32)
    H, W = D1raw.shape
    o = {k: np.zeros((H, W), np.float32) for k in
         ("D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2")}
    st = StageDump()
    for k, a in o.items():
        setattr(st, k, _ptr(a))
    if elas.param.subsampling:      # the injected maps are the half-resolution ones
        W, H = 2 * W, 2 * H
    d = (C.c_int32 * 3)(W, H, W)
    f = lib().jn_debug_postprocess
    f.argtypes = [_P, _P, _P, C.POINTER(C.c_int32), C.POINTER(StageDump)]
    _check(f(elas._h, _ptr(D1raw), _ptr(D2raw), d, C.byref(st)), "jn_debug_postprocess")
    return o


class Calibration:
    """K1,K2,D1,D2,R,T,XR,XT from the OpenCV YAML + the reprojection matrix Q."""

    def __init__(self, path=None):
        self.c = Calib()
        if path is not None:
            _check(lib().jn_calib_load_yaml(os.fsencode(path), C.byref(self.c)), "jn_calib_load_yaml")

    def set_q(self, cx, cy, f, tx):
        lib().jn_calib_set_q(C.byref(self.c), cx, cy, f, tx)

    def compose_cam_to_robot(self, phi_x, phi_y, phi_z, trans_x, trans_y, trans_z):
        """XR, XT from Euler angles + translation as the node's -m mode does (point_cloud.cpp:76-102, 305-311)."""
        f = lib().jn_calib_compose_cam_to_robot
        f.argtypes = [C.POINTER(Calib)] + [C.c_double] * 6
        _check(f(C.byref(self.c), phi_x, phi_y, phi_z, trans_x, trans_y, trans_z), "jn_calib_compose_cam_to_robot")

    def stereo_rectify(self, calib_w, calib_h, new_w=0, new_h=0, zero_disparity=True, alpha=0.0):
        """cv::stereoRectify as point_cloud.cpp:543-544 calls it, without OpenCV: sets Q, returns
        (R1, R2, P1, P2)."""
        R1 = np.zeros(9); R2 = np.zeros(9); P1 = np.zeros(12); P2 = np.zeros(12)
        f = lib().jn_calib_stereo_rectify
        f.argtypes = [C.POINTER(Calib), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, _P, _P]
        _check(f(C.byref(self.c), int(calib_w), int(calib_h), int(new_w), int(new_h), int(bool(zero_disparity)),
                 float(alpha), _ptr(R1), _ptr(R2), _ptr(P1), _ptr(P2)), "jn_calib_stereo_rectify")
        return R1.reshape(3, 3), R2.reshape(3, 3), P1.reshape(3, 4), P2.reshape(3, 4)

    def undistort_rectify_map(self, camera, R, P, w, h):
        """cv::initUndistortRectifyMap(K, D, R, P, (w, h), CV_32F) for camera 1 or 2 (point_cloud.cpp:553-554)."""
        K = np.ascontiguousarray(list(self.c.K1 if camera == 1 else self.c.K2), np.float64)
        D = np.ascontiguousarray(list(self.c.D1 if camera == 1 else self.c.D2), np.float64)
        R = np.ascontiguousarray(R, np.float64).reshape(9); P = np.ascontiguousarray(P, np.float64).reshape(12)
        mx = np.zeros((h, w), np.float32); my = np.zeros((h, w), np.float32)
        f = lib().jn_calib_init_undistort_rectify_map
        f.argtypes = [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P]
        _check(f(_ptr(K), _ptr(D), _ptr(R), _ptr(P), int(w), int(h), _ptr(mx), _ptr(my)), "jn_calib_init_undistort_rectify_map")
        return mx, my

    def set_q_matrix(self, Q):
        Q = np.asarray(Q, np.float64).reshape(16)
        for i in range(16):
            self.c.Q[i] = Q[i]
        self.c.has_q = 1

    def arrays(self):
        g = lambda a, s: np.array(list(a), np.float64).reshape(s)
        return {"K1": g(self.c.K1, (3, 3)), "K2": g(self.c.K2, (3, 3)), "D1": g(self.c.D1, (1, 5)),
                "D2": g(self.c.D2, (1, 5)), "R": g(self.c.R, (3, 3)), "T": g(self.c.T, (3,)),
                "XR": g(self.c.XR, (3, 3)), "XT": g(self.c.XT, (3, 1)), "Q": g(self.c.Q, (4, 4))}


class JpegDecoder:
    """cv::imdecode(..., CV_LOAD_IMAGE_GRAYSCALE) of point_cloud.cpp:436/478 for a batch of JPEG bitstreams,
    decoded by nvJPEG into device memory (jn_jpeg_decode_gray_batch)."""

    def __init__(self, device=0):
        l = lib()
        l.jn_jpeg_create.restype = _P
        l.jn_jpeg_create.argtypes = [C.c_int]
        l.jn_jpeg_destroy.argtypes = [_P]
        l.jn_jpeg_info.argtypes = [_P, _P, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.jn_jpeg_decode_gray_batch.argtypes = [_P, C.c_int, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_size_t, _P]
        self._h = l.jn_jpeg_create(int(device))
        if not self._h:
            raise JnError("jn_jpeg_create: " + last_error())

    def info(self, data):
        w = C.c_int32(0); h = C.c_int32(0)
        buf = np.frombuffer(bytes(data), np.uint8)
        _check(lib().jn_jpeg_info(self._h, _ptr(buf), buf.size, C.byref(w), C.byref(h)), "jn_jpeg_info")
        return w.value, h.value

    def decode_gray_batch(self, bitstreams, dst_ptr, width, height, dst_stride=None, frame_stride=None, stream=0):
        """bitstreams: list of bytes-like; dst_ptr: device address of len(bitstreams) frames."""
        bufs = [np.frombuffer(bytes(b), np.uint8) for b in bitstreams]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * n)(*[b.size for b in bufs])
        dst_stride = dst_stride or width
        frame_stride = frame_stride or dst_stride * height
        return _check(lib().jn_jpeg_decode_gray_batch(self._h, n, ptrs, lens, _P(int(dst_ptr)), width, height, dst_stride,
                                                      frame_stride, _P(stream) if stream else None), "jn_jpeg_decode_gray_batch")

    def close(self):
        if getattr(self, "_h", None):
            lib().jn_jpeg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ObstacleScan:
    """cacheDisparityValues + publishObstacleScan without ROS (point_cloud.cpp:104-147, 213-296)."""

    def __init__(self, calib, width, height, crop_offset_x=0, crop_offset_y=0, device=0):
        self.W, self.H = width, height
        self._h = lib().jn_scan_create(C.byref(calib.c), width, height, crop_offset_x, crop_offset_y, device)
        if not self._h:
            raise JnError("jn_scan_create: " + last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().jn_scan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fast_path(self):
        """include/jn_elas_debug.h: jn_scan_fast_path."""
        f = lib().jn_scan_fast_path
        f.restype = C.c_int
        f.argtypes = [C.c_void_p]
        return int(f(self._h))

    def gate_cache(self):
        out = np.zeros((self.H, self.W, 2), np.uint8)
        _check(lib().jn_scan_gate_cache(self._h, _ptr(out)), "jn_scan_gate_cache")
        return out

    def from_disparity(self, D, want_u8=False):
        D = _need(np.ascontiguousarray(D, np.float32), self.W * self.H, "disparity map")
        ranges = np.zeros(SCAN_BINS, np.float64)
        meta = ScanMeta()
        u8 = np.zeros((self.H, self.W), np.uint8) if want_u8 else None
        _check(lib().jn_scan_from_disparity(self._h, _ptr(D), _ptr(ranges), C.byref(meta), _ptr(u8)),
               "jn_scan_from_disparity")
        return ranges, meta, u8

    def from_disparity_batch(self, n, D, ranges, meta, dmap_u8=0, stream=0):
        """Device pointers (ints); asynchronous on `stream`."""
        return _check(lib().jn_scan_from_disparity_batch(self._h, int(n), _P(D), _P(ranges), _P(meta),
                                                         _P(dmap_u8) if dmap_u8 else None,
                                                         _P(stream) if stream else None),
                      "jn_scan_from_disparity_batch")

    def points(self, D):
        D = _need(np.ascontiguousarray(D, np.float32), self.W * self.H, "disparity map")
        pts = np.zeros((self.W * self.H, 3), np.float64)
        n = C.c_int32(0)
        ranges = np.zeros(SCAN_BINS, np.float64)
        meta = ScanMeta()
        _check(lib().jn_points_from_disparity(self._h, _ptr(D), _ptr(pts), C.byref(n), _ptr(ranges), C.byref(meta)),
               "jn_points_from_disparity")
        return pts[:n.value], ranges, meta

    def pointcloud(self, D, image):
        """sensor_msgs/PointCloud payload (point_cloud.cpp:351-383): (xyz float32 n x 3, rgb float32 n,
        ranges, meta).  image: H x W x 3 BGR or H x W grayscale (the reference's Vec3b-on-gray quirk)."""
        D = _need(np.ascontiguousarray(D, np.float32), self.W * self.H, "disparity map")
        image = np.ascontiguousarray(image, np.uint8)
        channels = 3 if image.ndim == 3 else 1
        if image.ndim not in (2, 3) or image.shape[0] < self.H or image.shape[1] < self.W or (image.ndim == 3 and image.shape[2] != 3):
            raise ValueError("image must be H x W (grayscale) or H x W x 3 (BGR) with H >= %d, W >= %d" % (self.H, self.W))
        xyz = np.zeros((self.W * self.H, 3), np.float32)
        rgb = np.zeros(self.W * self.H, np.float32)
        n = C.c_int32(0)
        ranges = np.zeros(SCAN_BINS, np.float64)
        meta = ScanMeta()
        _check(lib().jn_pointcloud_from_disparity(self._h, _ptr(D), _ptr(image), image.strides[0], channels, _ptr(xyz),
                                                  _ptr(rgb), C.byref(n), _ptr(ranges), C.byref(meta)),
               "jn_pointcloud_from_disparity")
        return xyz[:n.value], rgb[:n.value], ranges, meta


    def pointcloud_batch(self, n, D, xyz, counts, ranges, meta, rgb=0, image=0, image_stride=0, channels=1, stream=0):
        """The -g path for n frames, all arguments DEVICE addresses (ints); asynchronous on `stream`."""
        f = lib().jn_pointcloud_batch
        f.argtypes = [_P, C.c_int, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]
        P = lambda x: _P(int(x)) if x else None
        return _check(f(self._h, int(n), P(D), P(image), int(image_stride), int(channels), P(xyz), P(rgb), P(counts),
                        P(ranges), P(meta), P(stream)), "jn_pointcloud_batch")


def scan_compact(ranges):
    """LaserScan.ranges as the reference publishes them (finite bins, k = 89..0)."""
    ranges = _need(np.ascontiguousarray(ranges, np.float64), SCAN_BINS, "ranges")
    out = np.zeros(SCAN_BINS, np.float32)
    n = lib().jn_scan_compact(_ptr(ranges), _ptr(out))
    return out[:n]


NAV_STOP_IN_FRONT_MANUAL, NAV_OBSTACLE_AVOID, NAV_STOP_IN_FRONT = 0, 1, 2


class Navigate:
    """The scan's consumer without ROS: the `navigate` node's laserScanCallback / checkObstacle / chooseDirection
    (navigate.cpp:344-363, 101-153, 155-197).  Host code, no device needed."""

    def __init__(self):
        l = lib()
        l.jn_navigate_create.restype = _P
        l.jn_navigate_destroy.argtypes = [_P]
        l.jn_navigate_set_clearance.argtypes = [_P, C.c_double, C.c_double, C.c_int]
        l.jn_navigate_set_last_dir.argtypes = [_P, C.c_int]
        l.jn_navigate_last_dir.argtypes = [_P]
        l.jn_navigate_set_scan.argtypes = [_P, _P, C.c_int, C.c_double, C.c_double]
        l.jn_navigate_set_scan_bins.argtypes = [_P, _P, C.POINTER(ScanMeta)]
        l.jn_navigate_points.argtypes = [_P, _P, C.c_int]
        l.jn_navigate_check_obstacle.argtypes = [_P, _P]
        l.jn_navigate_choose_direction.argtypes = [_P]
        self._h = l.jn_navigate_create()
        if not self._h:
            raise JnError("jn_navigate_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().jn_navigate_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_clearance(self, clear_front, clear_side, laser_pt_thresh):
        """clear_front / clear_side / laser_pt_thresh (navigate.cpp:37-42; -c and -l of the node)."""
        lib().jn_navigate_set_clearance(self._h, clear_front, clear_side, int(laser_pt_thresh))

    def set_scan(self, ranges, angle_min, angle_max):
        """A LaserScan as point_cloud publishes it: compacted float32 ranges + angle_min / angle_max."""
        r = np.ascontiguousarray(ranges, np.float32)
        return _check(lib().jn_navigate_set_scan(self._h, _ptr(r), len(r), float(angle_min), float(angle_max)),
                      "jn_navigate_set_scan")

    def set_scan_bins(self, ranges, meta):
        """The 90-bin scan + ScanMeta of ObstacleScan.from_disparity directly."""
        r = np.ascontiguousarray(ranges, np.float64)
        if r.size != SCAN_BINS:
            raise ValueError("ranges must hold %d bins" % SCAN_BINS)
        return _check(lib().jn_navigate_set_scan_bins(self._h, _ptr(r), C.byref(meta)), "jn_navigate_set_scan_bins")

    def points(self):
        """laserPoints: n x 2 doubles (what visualizeLaserPoints publishes, navigate.cpp:77-98)."""
        n = lib().jn_navigate_points(self._h, None, 0)
        xy = np.zeros((max(n, 0), 2), np.float64)
        if n > 0:
            lib().jn_navigate_points(self._h, _ptr(xy), n)
        return xy

    def check_obstacle(self):
        """-> (isObstacle, (points in the safe box, laser points, closest distance, confidence of the vote))"""
        rep = (C.c_double * 4)()
        r = lib().jn_navigate_check_obstacle(self._h, rep)
        if r < 0:
            _check(r, "jn_navigate_check_obstacle")
        return r, (int(rep[0]), int(rep[1]), rep[2], rep[3])

    def choose_direction(self):
        """0 keep / 1 left / 2 right, with the reference's hysteresis on last_dir."""
        return lib().jn_navigate_choose_direction(self._h)

    def command(self, mode, side=0.0, front=0.0):
        """safeNavigate's velocity command for one mode (NAV_STOP_IN_FRONT_MANUAL / NAV_OBSTACLE_AVOID /
        NAV_STOP_IN_FRONT): (linear.x, angular.z) of the Twist the node publishes (navigate.cpp:302-342)."""
        f = lib().jn_navigate_command
        f.argtypes = [_P, C.c_int, C.c_double, C.c_double, _P]
        v = (C.c_double * 2)()
        _check(f(self._h, int(mode), float(side), float(front), v), "jn_navigate_command")
        return v[0], v[1]

    def set_max_forward_vel(self, v):
        f = lib().jn_navigate_set_max_forward_vel
        f.argtypes = [_P, C.c_float]
        f(self._h, float(v))

    @property
    def last_dir(self):
        return lib().jn_navigate_last_dir(self._h)

    @last_dir.setter
    def last_dir(self, d):
        lib().jn_navigate_set_last_dir(self._h, int(d))


class Rectifier:
    """cv::remap(frame, out, mapx, mapy, INTER_LINEAR) + ROI crop of one camera (point_cloud.cpp:440-442),
    with the CV_32FC1 map pair of cv::initUndistortRectifyMap (point_cloud.cpp:553-554)."""

    def __init__(self, mapx, mapy, device=0):
        mapx = np.ascontiguousarray(mapx, np.float32)
        mapy = np.ascontiguousarray(mapy, np.float32)
        if mapx.shape != mapy.shape or mapx.ndim != 2:
            raise ValueError("mapx and mapy must be 2-D arrays of the same shape")
        self.map_h, self.map_w = mapx.shape
        self._h = lib().jn_rectify_create(_ptr(mapx), _ptr(mapy), self.map_w, self.map_h, int(device))
        if not self._h:
            raise JnError("jn_rectify_create: " + last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().jn_rectify_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def remap_batch(self, src, n, src_w, src_h, src_stride, dst, dst_stride, roi=None, stream=0):
        """Device pointers (ints); asynchronous on `stream`."""
        r = (C.c_int32 * 4)(*[int(x) for x in roi]) if roi is not None else None
        return _check(lib().jn_rectify_batch(self._h, int(n), _P(src), int(src_w), int(src_h), int(src_stride), r,
                                             _P(dst), int(dst_stride), _P(stream) if stream else None),
                      "jn_rectify_batch")
