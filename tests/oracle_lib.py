"""ctypes access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

  ref  = oracle/_ref/libelas_ref.so  (unmodified reference, built by `make -C oracle ref`)
  port = oracle/libelas_port.so      (plain-C restatement, built by `make -C oracle port`)

Both export the same functions (prefix ref_ / port_), see oracle/oracle_abi.h.
"""
import ctypes as C
import math
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libelas_ref.so")
PORT_SO = os.path.join(ROOT, "oracle", "libelas_port.so")


class Params(C.Structure):
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


def robotics(disp_max=255, **kw):
    """Elas::parameters(ROBOTICS) (elas.h:92-115) + overrides."""
    p = Params(0, disp_max, 0.85, 10, 5, 5, 5, 5, 0, 20, 0.02, 3.0, 1.0, 2.0,
               1, 2, 1.0, 200, 3, 0, 1, 1, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def middlebury(disp_max=255, **kw):
    """Elas::parameters(MIDDLEBURY) (elas.h:119-143) + overrides."""
    p = Params(0, disp_max, 0.95, 10, 5, 5, 5, 5, 1, 20, 0.02, 5.0, 1.0, 3.0,
               0, 2, 1.0, 200, 5000, 1, 0, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


_P = C.c_void_p


class Stages(C.Structure):
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


def lattice_dims(W, H, step=5):
    return (W + step - 1) // step, (H + step - 1) // step


def effective_step(p):
    """elas.cpp:379-381: with subsampling only even rows carry descriptors, so an odd step is bumped."""
    s = p.candidate_stepsize
    return s + s % 2 if p.subsampling else s


def map_dims(p, W, H):
    """(rows, cols) of the disparity maps (elas.cpp:58-63 of main.cpp / elas.h:120-124)."""
    return (H // 2, W // 2) if p.subsampling else (H, W)


def grid_dims(W, H, gs=20):
    return int(math.ceil(np.float32(W) / np.float32(gs))), int(math.ceil(np.float32(H) / np.float32(gs)))


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def aligned(shape, dtype, align=64):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


class Oracle:
    """One of the two CPU checkers."""

    def __init__(self, kind):
        self.kind = kind
        path = REF_SO if kind == "ref" else PORT_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.pfx = "ref_" if kind == "ref" else "port_"
        if kind == "ref":
            self.lib.ref_set_deterministic_heap(1)

    def fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def has(self, name):
        return hasattr(self.lib, self.pfx + name)

    # ---- full pipeline -------------------------------------------------
    def process(self, p, I1, I2):
        H, W = I1.shape
        I1 = np.ascontiguousarray(I1)
        I2 = np.ascontiguousarray(I2)
        Hd, Wd = map_dims(p, W, H)
        D1 = np.zeros((Hd, Wd), np.float32)
        D2 = np.zeros((Hd, Wd), np.float32)
        dims = (C.c_int32 * 3)(W, H, W)
        self.fn("elas_process")(C.byref(p), _ptr(I1), _ptr(I2), _ptr(D1), _ptr(D2), dims)
        return D1, D2

    def stages(self, p, I1, I2, want_desc=True, want_grid=True):
        """Runs the whole pipeline and returns a dict of every intermediate."""
        H, W = I1.shape
        I1 = np.ascontiguousarray(I1)
        I2 = np.ascontiguousarray(I2)
        Wc, Hc = lattice_dims(W, H, effective_step(p))
        gw, gh = grid_dims(W, H, p.grid_size)
        Hd, Wd = map_dims(p, W, H)
        cap_s = Wc * Hc + 8
        cap_t = 2 * cap_s + 8
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
            o[k] = np.zeros((H, W), np.float32)
        st = Stages()
        for k, a in o.items():
            setattr(st, k, _ptr(a))
        st.cap_support = cap_s
        st.cap_tri = cap_t
        dims = (C.c_int32 * 3)(W, H, W)
        f = self.fn("elas_stages")
        f.restype = C.c_int
        rc = f(C.byref(p), _ptr(I1), _ptr(I2), dims, C.byref(st))
        o["rc"] = rc
        o["n_support"] = st.n_support
        o["support"] = o["support"][:st.n_support]
        if rc == 0:
            o["tri1"] = o["tri1"][:st.n_tri1]
            o["tri2"] = o["tri2"][:st.n_tri2]
            o["planes1"] = o["planes1"][:st.n_tri1]
            o["planes2"] = o["planes2"][:st.n_tri2]
        o["dense_evals"] = st.dense_evals
        o["dense_pixels"] = st.dense_pixels
        if p.subsampling:   # the maps are (H/2) x (W/2), stored contiguously at the front of each buffer
            for k in ("D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap",
                      "D1_mean", "D2_mean", "D1", "D2"):
                o[k] = o[k].reshape(-1)[:Hd * Wd].reshape(Hd, Wd).copy()
        return o

    # ---- single stages with injected inputs ------------------------------
    def descriptor(self, I):
        H, W = I.shape
        I = np.ascontiguousarray(I)
        out = np.zeros((H, W, 16), np.uint8)
        self.fn("descriptor")(_ptr(I), W, H, W, _ptr(out))
        return out

    def triangulate(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = xy.shape[0]
        tri = np.zeros((2 * n + 8, 3), np.int32)
        f = self.fn("triangulate")
        f.restype = C.c_int
        nt = f(_ptr(xy), n, _ptr(tri), tri.shape[0])
        return tri[:nt].copy()

    def filter_dcan(self, p, dcan):
        d = np.ascontiguousarray(dcan, np.int16).copy()
        Hc, Wc = d.shape
        inc = np.zeros_like(d)
        self.fn("filter_dcan")(C.byref(p), _ptr(d), Wc, Hc, _ptr(inc))
        return inc, d

    def planes(self, p, support, tri, right_image):
        support = np.ascontiguousarray(support, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        out = np.zeros((tri.shape[0], 6), np.float32)
        self.fn("planes")(C.byref(p), _ptr(support), support.shape[0], _ptr(tri), tri.shape[0],
                          int(right_image), _ptr(out))
        return out

    def grid(self, p, W, H, support, right_image):
        support = np.ascontiguousarray(support, np.int32)
        gw, gh = grid_dims(W, H, p.grid_size)
        out = np.zeros((gh, gw, p.disp_max + 2), np.int32)
        self.fn("grid")(C.byref(p), W, H, _ptr(support), support.shape[0], int(right_image), _ptr(out))
        return out

    def dense(self, p, desc1, desc2, support, tri, planes, grid, right_image):
        H, W, _ = desc1.shape
        d1 = aligned((H, W, 16), np.uint8); d1[...] = desc1
        d2 = aligned((H, W, 16), np.uint8); d2[...] = desc2
        support = np.ascontiguousarray(support, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        planes = np.ascontiguousarray(planes, np.float32)
        grid = np.ascontiguousarray(grid, np.int32)
        D = np.zeros((H, W), np.float32)
        self.fn("dense")(C.byref(p), W, H, _ptr(d1), _ptr(d2), _ptr(support), support.shape[0],
                         _ptr(tri), _ptr(planes), tri.shape[0], _ptr(grid), int(right_image), _ptr(D))
        return D

    def postprocess(self, p, D1, D2):
        H, W = D1.shape
        a = np.ascontiguousarray(D1, np.float32).copy()
        b = np.ascontiguousarray(D2, np.float32).copy()
        o = {k: np.zeros((H, W), np.float32) for k in
             ("D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2")}
        st = Stages()
        for k, arr in o.items():
            setattr(st, k, _ptr(arr))
        if p.subsampling:           # D1/D2 are the half-resolution maps; the oracle takes the image size
            W, H = 2 * W, 2 * H
        self.fn("postprocess")(C.byref(p), W, H, _ptr(a), _ptr(b), C.byref(st))
        return o


def load(kind):
    try:
        return Oracle(kind)
    except (FileNotFoundError, OSError):
        return None


def port_remap(src, mapx, mapy, roi=None):
    """oracle/remap_port.c: cv::remap(INTER_LINEAR, constant border 0) + ROI crop on the CPU."""
    lib = C.CDLL(PORT_SO)
    src = np.ascontiguousarray(src, np.uint8)
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    H, W = mapx.shape
    x0, y0, w, h = roi if roi is not None else (0, 0, W, H)
    dst = np.zeros((h, w), np.uint8)
    lib.port_remap_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(mapx), _ptr(mapy), W,
                      x0, y0, w, h, _ptr(dst), dst.strides[0])
    return dst


def remap_golden_cases():
    """(name, src, mapx, mapy, dst) of tests/golden/remap_cv2.npz (outputs of cv2.remap)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "remap_cv2.npz"))
    for k in ("calib0", "calib1", "rand"):
        if k + "_src" in z.files:
            src = z[k + "_src"]
        else:
            sd, h, w = [int(x) for x in z[k + "_src_seed"]]
            src = np.random.default_rng(sd).integers(0, 256, (h, w), dtype=np.uint8)
        yield k, src, z[k + "_mapx"], z[k + "_mapy"], z[k + "_dst"]
