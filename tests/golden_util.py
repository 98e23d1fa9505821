"""Helpers to compare a stage dump with a golden fixture (tests/golden/*.npz)."""
import os
import numpy as np
import oracle_lib as ol

GOLD = os.path.join(ol.ROOT, "tests", "golden")
EXACT = ["dcan_raw", "dcan_incon", "dcan_final", "support", "tri1", "tri2", "planes1", "planes2",
         "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2"]


def load(name):
    z = np.load(os.path.join(GOLD, name))
    p = ol.Params.from_buffer_copy(z["params"].tobytes())
    return z, p


def check(stages, z, H):
    """Asserts that every stage of `stages` equals the golden fixture `z` bit for bit."""
    for k in EXACT:
        a, b = np.asarray(stages[k]), z[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.array_equal(a, b), "%s: %d mismatches" % (k, int((a != b).sum()))
    for s in ("1", "2"):
        d = np.asarray(stages["desc" + s])
        assert np.array_equal(d.astype(np.uint32).sum(axis=(1, 2)).astype(np.uint32), z["desc%s_rowsum" % s])
        g = np.asarray(stages["grid" + s])
        assert np.array_equal(g[..., 0], z["grid%s_count" % s])
        assert np.array_equal(g[..., 1:].sum(axis=2).astype(np.int32), z["grid%s_sum" % s])
    assert np.array_equal(np.asarray(stages["desc1"])[H // 2 - 2:H // 2 + 3], z["desc1_band"])
