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
