"""Randomised single-stage parity tests (GPU): inputs the synthetic stereo scenes never
produce are injected straight into one stage and compared with the oracle.
  - support filter: noisy candidate images where the in-place, scan-ordered inconsistency
    filter cascades (several frontier rounds on the GPU)
  - Delaunay: lattice subsets, random integers, duplicates, collinear sets
  - post-processing: maps full of speckles, gaps of every width and disparity jumps
All comparisons are bit-exact."""
import numpy as np
import pytest
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def random_dcan(rng, Hc, Wc, mode):
    yy, xx = np.mgrid[0:Hc, 0:Wc]
    base = (20 + 0.3 * yy + 0.1 * xx).astype(np.int32)
    if mode == 0:      # sparse noise: almost everything inconsistent
        d = rng.integers(0, 200, (Hc, Wc)); keep = rng.random((Hc, Wc)) < 0.3
    elif mode == 1:    # smooth surface with dropouts + outliers: cascades at the dropout borders
        d = base + rng.integers(-1, 2, (Hc, Wc)); keep = rng.random((Hc, Wc)) < 0.6
        out = rng.random((Hc, Wc)) < 0.05
        d = np.where(out, rng.integers(0, 200, (Hc, Wc)), d)
    elif mode == 2:    # two layers interleaved: support counts hover around the threshold
        d = np.where(rng.random((Hc, Wc)) < 0.5, base, base + 40); keep = rng.random((Hc, Wc)) < 0.25
    elif mode == 4:    # chains of points that each have exactly incon_min_support supporters: the
        # first two of a chain fall short, and every removal pulls the next point below the
        # threshold -> one frontier round per chain link (the in-place cascade of H3)
        d = np.full((Hc, Wc), 50); keep = np.zeros((Hc, Wc), bool)
        for r in range(8, Hc - 8, 14):
            keep[r, 6:Wc - 6:2] = True                       # horizontal chain, spacing 2
        for c in range(10, Wc - 10, 17):
            keep[3:Hc // 2:2, c] |= rng.random() < 0.5       # some vertical chains crossing them
        d = d + (yy // 14) * 7                               # chains do not support each other
    else:              # dense and constant: the redundancy passes do all the work
        d = np.full((Hc, Wc), 33); keep = rng.random((Hc, Wc)) < 0.95
    out = np.where(keep, d, -1).astype(np.int16)
    out[0, :] = 0; out[:, 0] = 0      # the reference's calloc'd row/column 0 (H3)
    return out


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_support_filter_random(jn, oracle, mode):
    rng = np.random.default_rng(100 + mode)
    max_rounds = 0
    for W, H in ((320, 240), (333, 251), (640, 480)):
        Wc, Hc = ol.lattice_dims(W, H)
        e = jn.Elas(jn.parameters(jn.ROBOTICS))
        for rep in range(3):
            dcan = random_dcan(rng, Hc, Wc, mode)
            inc_ref, fin_ref = oracle.filter_dcan(ol.robotics(), dcan)
            inc, fin, sup, rounds = jn.debug_support_filter(e, dcan, W, H)
            max_rounds = max(max_rounds, rounds)
            assert np.array_equal(inc, inc_ref), (mode, W, rep, int((inc != inc_ref).sum()))
            assert np.array_equal(fin, fin_ref), (mode, W, rep, int((fin != fin_ref).sum()))
            # emission order: u outer, v inner, lattice row/column 0 excluded (elas.cpp:426-431)
            vv, uu = np.nonzero(fin_ref[1:, 1:].T >= 0)[::-1]
            exp = np.stack([(uu + 1) * 5, (vv + 1) * 5, fin_ref[vv + 1, uu + 1]], 1) if len(uu) else np.zeros((0, 3))
            order = np.lexsort((exp[:, 1], exp[:, 0])) if len(exp) else []
            assert np.array_equal(sup, exp[order].astype(np.int32))
        e.close()
    if mode == 4:
        assert max_rounds >= 10, "the cascade (frontier propagation) was not exercised: %d rounds" % max_rounds


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_delaunay_random(jn, oracle, mode):
    rng = np.random.default_rng(200 + mode)
    e = jn.Elas(jn.parameters(jn.ROBOTICS))
    for it in range(25):
        n = int(rng.integers(3, 1500))
        if mode == 0:
            g = int(rng.integers(2, 60)); pts = rng.integers(1, g + 1, size=(n, 2)) * 5
        elif mode == 1:
            pts = rng.integers(0, 2000, size=(n, 2))
        elif mode == 2:
            pts = rng.integers(1, 7, size=(min(n, 80), 2)) * 5                # many duplicates
        elif mode == 3:
            pts = np.stack([rng.integers(1, 60, size=min(n, 200)) * 5, np.full(min(n, 200), 10)], 1)   # collinear
        else:
            u = rng.integers(1, 380, size=n) * 5; v = rng.integers(1, 240, size=n) * 5
            pts = np.stack([u - rng.integers(0, 60, size=n) + 64, v], 1)      # right-image like (biased >= 0)
        if mode in (0, 1, 4):
            pts = np.unique(pts, axis=0)
            rng.shuffle(pts)
        if len(pts) < 3:
            continue
        ref = oracle.triangulate(pts)
        got = jn.debug_triangulate(e, pts, 1920, 1200)
        assert got.shape == ref.shape, (mode, it, got.shape, ref.shape)
        # mode 2, coincident points: Triangle keeps the copy its randomised quicksort puts first; the
        # kernel replays that quicksort (csrc/vertexsort.cuh), so the vertex ids are equal too
        assert np.array_equal(got, ref), (mode, it)
    e.close()


@pytest.mark.parametrize("smem_max", [6, 40, 333, 2000])
def test_delaunay_subtree_tiling(jn, oracle, smem_max):
    """Point sets above the shared-memory table limit are cut into subtrees of the D&C tree that are
    built in shared memory one after the other, the top merges run on the global table.  Force small
    limits so that 1 to 9 levels end up on the global side; then one real case: 15 000 points."""
    rng = np.random.default_rng(900 + smem_max)
    e = jn.Elas(jn.parameters(jn.ROBOTICS))
    try:
        jn.lib().jn_debug_delaunay_limits(-1, smem_max)
        for it in range(8):
            n = int(rng.integers(smem_max + 1, 6000))
            if it % 2:
                pts = np.stack([rng.integers(1, 380, size=n) * 5 - rng.integers(0, 60, size=n) + 64,
                                rng.integers(1, 240, size=n) * 5], 1)
            else:
                pts = rng.integers(1, 120, size=(n, 2)) * 5
            pts = np.unique(pts, axis=0)
            rng.shuffle(pts)
            ref = oracle.triangulate(pts)
            got = jn.debug_triangulate(e, pts, 1920, 1200)
            assert got.shape == ref.shape and np.array_equal(got, ref), (smem_max, it, len(pts))
    finally:
        jn.lib().jn_debug_delaunay_limits(-1, -1)
    e.close()


def test_delaunay_15000_points(jn, oracle):
    """Above 8 192 points: two subtrees in shared memory, the root merge on the global table."""
    rng = np.random.default_rng(77)
    u = rng.integers(1, 384, size=40000) * 5; v = rng.integers(1, 240, size=40000) * 5
    pts = np.unique(np.stack([u, v], 1), axis=0)
    rng.shuffle(pts)
    pts = pts[:15000]
    e = jn.Elas(jn.parameters(jn.ROBOTICS))
    ref = oracle.triangulate(pts)
    got = jn.debug_triangulate(e, pts, 1920, 1200)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    e.close()


def random_disparity_maps(rng, H, W, mode):
    yy, xx = np.mgrid[0:H, 0:W]
    surf = np.floor(10 + 0.08 * yy + 0.01 * xx)
    D1 = surf.copy()
    if mode == 0:      # speckles: random blobs of other disparities, many below 200 px
        for _ in range(60):
            y, x = rng.integers(0, H), rng.integers(0, W); r = rng.integers(1, 12)
            D1[max(0, y - r):y + r, max(0, x - r):x + r] = rng.integers(0, 60)
        D1[rng.random((H, W)) < 0.05] = -1
    elif mode == 1:    # gaps of every width along rows and columns
        for _ in range(150):
            y, x = rng.integers(0, H), rng.integers(0, W); L = rng.integers(1, 9)
            if rng.random() < 0.5: D1[y, x:x + L] = -10
            else: D1[y:y + L, x] = -10
        D1 += rng.integers(0, 2, (H, W))
    else:              # salt and pepper on a two-level surface: stresses mean / median filters
        D1 = np.where(xx > W // 2, surf + 7, surf) + rng.integers(-1, 2, (H, W))
        D1[rng.random((H, W)) < 0.15] = -10
    D1 = D1.astype(np.float32)
    # a right map that is mostly consistent with D1: D2(u - d, v) = d
    D2 = np.full((H, W), -10, np.float32)
    ys, xs = np.nonzero(D1 >= 0)
    xr = xs - D1[ys, xs].astype(np.int64)
    ok = xr >= 0
    D2[ys[ok], xr[ok]] = D1[ys[ok], xs[ok]]
    D2[rng.random((H, W)) < 0.03] = rng.integers(0, 50)
    return D1, D2


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kw", [{}, {"filter_median": 1, "postprocess_only_left": 0},
                                {"ipol_gap_width": 30, "speckle_size": 40, "postprocess_only_left": 0},
                                {"speckle_sim_threshold": 2.0, "lr_threshold": 1, "filter_adaptive_mean": 0}])
def test_postprocess_random(jn, oracle, mode, kw):
    rng = np.random.default_rng(300 + mode)
    H, W = 150, 211
    D1, D2 = random_disparity_maps(rng, H, W, mode)
    ref = oracle.postprocess(ol.robotics(64, **kw), D1, D2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=64, **kw))
    got = jn.debug_postprocess(e, D1, D2)
    for k in ("D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2"):
        assert np.array_equal(got[k], ref[k]), "%s: %d pixels differ" % (k, int((got[k] != ref[k]).sum()))
    e.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kw", [{}, {"filter_median": 1, "postprocess_only_left": 0},
                                {"ipol_gap_width": 30, "speckle_size": 40, "postprocess_only_left": 0}])
def test_postprocess_random_subsampled(jn, oracle, mode, kw):
    """Half-resolution branches: d/2 warp in the L/R check, rescaled segment and gap limits, 4-tap mean."""
    rng = np.random.default_rng(400 + mode)
    H, W = 150, 211
    D1, D2 = random_disparity_maps(rng, H, W, mode)
    ref = oracle.postprocess(ol.robotics(64, subsampling=1, **kw), D1, D2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=64, subsampling=1, **kw))
    got = jn.debug_postprocess(e, D1, D2)
    for k in ("D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap", "D1_mean", "D2_mean", "D1", "D2"):
        assert np.array_equal(got[k], ref[k]), "%s: %d pixels differ" % (k, int((got[k] != ref[k]).sum()))
    e.close()
