"""Synthetic random-dot stereo pairs with known ground truth.

The reference ships no images (its data/ directory is git-ignored), so every
test and benchmark input is generated here.  Left image: iid uniform u8 dots.
Ground truth: a slanted ground plane d(u,v) = floor(0.10*dmax + 0.50*dmax*v/H)
with one fronto-parallel box at floor(0.6*dmax).  Right image painted far to
near with I2(u-d, v) = I1(u, v); right pixels nobody painted are independent
noise.
"""
import numpy as np


def ground_truth(W, H, disp_max):
    v = np.arange(H, dtype=np.float64)[:, None]
    d = np.floor(0.10 * disp_max + 0.50 * disp_max * v / H).astype(np.int32)
    d = np.broadcast_to(d, (H, W)).copy()
    u = np.arange(W)[None, :]
    vv = np.arange(H)[:, None]
    box = (u > W // 3) & (u < 2 * W // 3) & (vv > H // 4) & (vv < 2 * H // 3)
    d[box] = int(np.floor(0.6 * disp_max))
    return d


def synth_pair(W, H, disp_max, seed):
    """Returns (I1, I2, gt) : uint8 HxW, uint8 HxW, int32 HxW."""
    rng = np.random.Generator(np.random.PCG64(seed))
    I1 = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    rng2 = np.random.Generator(np.random.PCG64(seed + 1))
    I2 = rng2.integers(0, 256, size=(H, W), dtype=np.uint8)
    gt = ground_truth(W, H, disp_max)
    # paint far -> near so nearer surfaces occlude farther ones
    uu = np.broadcast_to(np.arange(W)[None, :], (H, W))
    vv = np.broadcast_to(np.arange(H)[:, None], (H, W))
    for d in np.unique(gt):
        m = (gt == d) & (uu - d >= 0)
        I2[vv[m], uu[m] - d] = I1[m]
    return I1, I2, gt


def synth_batch(W, H, disp_max, seeds):
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    for i, s in enumerate(seeds):
        L[i], R[i], _ = synth_pair(W, H, disp_max, s)
    return L, R


# ----------------------------------------------------------------------------------------------
# Second scene family: band-limited texture, sub-pixel disparities slanted in u and v,
# occluding boxes and a textureless patch.  Random dots with integer ground truth are the
# easiest input ELAS can get; this one produces more (and less regular) support points, cells
# with many disparity candidates, failed matches and real occlusions.

def _pink_texture(rng, H, W, alpha=1.2):
    """1/f^alpha noise, mean 128, clipped to [0,255] (float64, H x W)."""
    white = rng.standard_normal((H, W))
    F = np.fft.rfft2(white)
    fy = np.fft.fftfreq(H)[:, None]
    fx = np.fft.rfftfreq(W)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    F *= 1.0 / np.power(np.maximum(f, 1.0 / 64.0), alpha)   # flat below 1/64 cycles/px: no huge blobs
    F[0, 0] = 0.0
    t = np.fft.irfft2(F, s=(H, W))
    t *= 52.0 / t.std()
    return np.clip(t + 128.0, 0.0, 255.0)


def textured_layers(W, H, disp_max, seed):
    """Layers far -> near: (a, b, c, u0, u1, v0, v1) with d(u,v) = a*u + b*v + c inside the box
    [u0,u1) x [v0,v1) of the LEFT image.  Layer 0 is the background (slanted in u and v)."""
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    dm = float(disp_max)
    layers = [(0.25 * dm / W, 0.46 * dm / H, 0.07 * dm + 0.3, 0.0, float(W), 0.0, float(H))]
    for k in range(3):
        bw = W * rng.uniform(0.12, 0.22)
        bh = H * rng.uniform(0.15, 0.30)
        u0 = rng.uniform(0.05 * W, 0.95 * W - bw)
        v0 = rng.uniform(0.05 * H, 0.80 * H - bh)
        # nearer than the background everywhere inside the box
        a0, b0, c0 = layers[0][:3]
        d_bg = a0 * (u0 + bw) + b0 * (v0 + bh) + c0
        d0 = min(d_bg + dm * rng.uniform(0.06, 0.18), 0.92 * dm)
        a = rng.uniform(-0.03, 0.03) * dm / W
        layers.append((a, 0.0, d0 - a * (u0 + 0.5 * bw) + rng.uniform(0, 1), u0, u0 + bw, v0, v0 + bh))
    return layers


def textured_pair(W, H, disp_max, seed):
    """Returns (I1, I2, gt): uint8 HxW, uint8 HxW, float32 HxW (sub-pixel left disparity)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    layers = textured_layers(W, H, disp_max, seed)
    pad = 64
    tex = _pink_texture(rng, H, (W + pad) * len(layers))
    # textureless patch on the background
    pw, ph = int(0.10 * W), int(0.12 * H)
    pu, pv = int(rng.uniform(0.05, 0.80) * W), int(rng.uniform(0.55, 0.85) * H)
    tex[pv:pv + ph, pu:pu + pw] = 117.0
    I1 = np.zeros((H, W), np.float64)
    I2 = _pink_texture(rng, H, W, 0.6)        # what no left pixel explains (image border, u - d < 0)
    gt = np.zeros((H, W), np.float32)
    uu = np.arange(W, dtype=np.float64)[None, :]
    vv = np.arange(H, dtype=np.float64)[:, None]
    rows = np.broadcast_to(np.arange(H)[:, None], (H, W))
    for k, (a, b, c, u0, u1, v0, v1) in enumerate(layers):
        T = tex[:, k * (W + pad):(k + 1) * (W + pad)]
        inside_v = (vv >= v0) & (vv < v1)
        m = inside_v & (uu >= u0) & (uu < u1)
        I1[m] = np.broadcast_to(T[:, :W], (H, W))[m]
        gt[m] = np.broadcast_to(a * uu + b * vv + c, (H, W))[m].astype(np.float32)
        # right image: the left coordinate u that lands on right pixel ur:  u - d(u,v) = ur
        ul = (uu + b * vv + c) / (1.0 - a)
        mr = inside_v & (ul >= u0) & (ul < u1) & (ul >= 0) & (ul <= W - 1)
        i0 = np.floor(ul).astype(np.int64)
        fr = ul - i0
        i0c = np.clip(i0, 0, W + pad - 2)
        s = T[rows, i0c] * (1.0 - fr) + T[rows, i0c + 1] * fr
        I2[mr] = s[mr]
    n1 = rng.integers(-2, 3, size=(H, W))
    n2 = rng.integers(-2, 3, size=(H, W))
    I1 = np.clip(np.rint(I1) + n1, 0, 255).astype(np.uint8)
    I2 = np.clip(np.rint(I2) + n2, 0, 255).astype(np.uint8)
    return I1, I2, gt


SCENES = {"random_dot": synth_pair, "textured": textured_pair}


def scene_batch(scene, W, H, disp_max, seeds):
    fn = SCENES[scene]
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    for i, s in enumerate(seeds):
        L[i], R[i], _ = fn(W, H, disp_max, s)
    return L, R
