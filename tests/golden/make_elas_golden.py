"""Generates the golden ELAS fixtures from the UNMODIFIED reference (oracle/_ref, built from
/root/reference).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_elas_golden.py

Fixtures (small, compressed):
  elas_robotics_160x120.npz   ROBOTICS preset, disp_max 48, postprocess_only_left
  elas_c5_200x150.npz         ROBOTICS + filter_median + both images post-processed, disp_max 64
  elas_sub_240x180.npz        ROBOTICS + subsampling (half-resolution maps), disp_max 48
  delaunay_cases.npz          point sets (lattice / cocircular / collinear / duplicates) with the
                              triangle lists the real Triangle ("zQB") returns for them
"""
import importlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
synth = importlib.import_module("jackal-navigation_b200.synth")

ref = ol.load("ref")
assert ref is not None, "build oracle/_ref first"

KEYS = ["dcan_raw", "dcan_incon", "dcan_final", "support", "tri1", "tri2", "planes1", "planes2",
        "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2"]


def dump(name, W, H, dm, seed, **kw):
    I1, I2, gt = synth.synth_pair(W, H, dm, seed)
    p = ol.robotics(dm, **kw)
    o = ref.stages(p, I1, I2)
    assert o["rc"] == 0
    out = {k: o[k] for k in KEYS}
    # descriptors: keep a checksum per row + one full row band (they are 16x the image)
    out["desc1_rowsum"] = o["desc1"].astype(np.uint32).sum(axis=(1, 2)).astype(np.uint32)
    out["desc2_rowsum"] = o["desc2"].astype(np.uint32).sum(axis=(1, 2)).astype(np.uint32)
    out["desc1_band"] = o["desc1"][H // 2 - 2:H // 2 + 3].copy()
    out["grid1_count"] = o["grid1"][..., 0].copy()
    out["grid2_count"] = o["grid2"][..., 0].copy()
    out["grid1_sum"] = o["grid1"][..., 1:].sum(axis=2).astype(np.int32)
    out["grid2_sum"] = o["grid2"][..., 1:].sum(axis=2).astype(np.int32)
    out["I1"] = I1; out["I2"] = I2
    out["params"] = np.frombuffer(bytes(p), np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "support", o["n_support"], "tri", len(o["tri1"]), len(o["tri2"]),
          "valid %.3f" % (o["D1"] >= 0).mean(), os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


dump("elas_robotics_160x120.npz", 160, 120, 48, 3)
dump("elas_c5_200x150.npz", 200, 150, 64, 11, filter_median=1, postprocess_only_left=0)
dump("elas_sub_240x180.npz", 240, 180, 48, 21, subsampling=1, postprocess_only_left=0)

rng = np.random.default_rng(42)
cases = {}
def add(name, pts):
    pts = np.asarray(pts, np.int32)
    cases["pts_" + name] = pts
    cases["tri_" + name] = ref.triangulate(pts)
g = np.stack(np.meshgrid(np.arange(1, 9) * 5, np.arange(1, 7) * 5, indexing="ij"), -1).reshape(-1, 2)
add("full_lattice", g)
add("lattice_subset", g[rng.permutation(len(g))[:30]])
add("random", np.unique(rng.integers(0, 500, (120, 2)), axis=0))
add("collinear", np.stack([np.arange(3, 20) * 5, np.full(17, 35)], 1))
add("with_duplicates", rng.integers(0, 6, (60, 2)) * 5)
u = rng.integers(1, 100, 200) * 5; v = rng.integers(1, 60, 200) * 5; d = rng.integers(0, 40, 200)
add("right_image_like", np.unique(np.stack([u - d, v], 1), axis=0))
add("three", [[5, 5], [10, 5], [5, 10]])
add("four_cocircular", [[5, 5], [10, 5], [5, 10], [10, 10]])
np.savez_compressed(os.path.join(HERE, "delaunay_cases.npz"), **cases)
print("delaunay_cases.npz", len(cases) // 2, "cases")
