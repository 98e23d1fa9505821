"""The reference-shaped calls of the C ABI on the GPU: a fresh Elas per frame (point_cloud.cpp:416-419)
without allocations, D2 = NULL, and the host-buffer batch entry point (image pairs in, scans out)."""
import ctypes as C
import numpy as np
import pytest
import oracle_lib as ol
import scan_lib

pytestmark = pytest.mark.gpu


def test_fresh_elas_per_frame_reuses_device_resources(jn, oracle, synth):
    """point_cloud.cpp:416-419 constructs an Elas for every frame; the handles hand their device
    resources to each other through the library's cache (no cudaMalloc per frame), results unchanged."""
    import torch
    W, H, dm = 320, 240, 64
    frames = [synth.synth_pair(W, H, dm, s)[:2] for s in (41, 42, 43)]
    refs = [oracle.process(ol.robotics(dm), a, b) for a, b in frames]
    jn.lib().jn_cache_clear()
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm)); D1 = np.zeros((H, W), np.float32); D2 = D1.copy()
    e.process(frames[0][0], frames[0][1], D1, D2, (W, H, W)); e.close()          # warm: allocates
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for rep in range(3):
        for (a, b), (R1, R2) in zip(frames, refs):
            e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
            D1 = np.zeros((H, W), np.float32); D2 = np.zeros((H, W), np.float32)
            assert e.process(a, b, D1, D2, (W, H, W)) == 0
            assert np.array_equal(D1, R1) and np.array_equal(D2, R2)
            e.close()
    assert torch.cuda.mem_get_info()[0] == free0          # nothing allocated or freed in between
    jn.lib().jn_cache_clear()
    assert torch.cuda.mem_get_info()[0] > free0


def test_process_without_right_map(jn, oracle, synth):
    W, H, dm = 333, 251, 100
    I1, I2, _ = synth.textured_pair(W, H, dm, 5)
    R1, _ = oracle.process(ol.robotics(dm), I1, I2)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    D1 = np.zeros((H, W), np.float32)
    assert e.process(I1, I2, D1, None, (W, H, W)) == 0
    assert np.array_equal(D1, R1)
    with pytest.raises(ValueError):
        e.process(I1[:100], I2, D1, None, (W, H, W))      # image smaller than dims says
    e.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_host_batch_equals_per_frame_calls(jn, oracle, synth, pinned):
    """jn_stereo_scan_submit / _wait: three overlapping submissions of host buffers (one textureless
    frame among them) give per-frame exactly what Elas::process + the scan give."""
    import torch
    W, H, dm, n = 320, 240, 64, 4
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(scan_lib.fixtures()["Q"]["320x180"])
    sc = jn.ObstacleScan(cal, W, H)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    sets = []
    for call in range(3):
        L, R = synth.scene_batch("textured" if call == 1 else "random_dot", W, H, dm, [60 + 10 * call + i for i in range(n)])
        if call == 2:
            L[1] = 7; R[1] = 7
        sets.append(dict(L=mk(torch.from_numpy(L)), R=mk(torch.from_numpy(R)),
                         D1=mk(torch.full((n, H, W), 5.0)), st=mk(torch.full((n,), -9, dtype=torch.int32)),
                         ranges=mk(torch.zeros((n, 90), dtype=torch.float64)), meta=mk(torch.zeros((n, 5), dtype=torch.float64)),
                         u8=mk(torch.zeros((n, H, W), dtype=torch.uint8))))
    for s in sets:
        e.stereo_scan_submit(sc, n, s["L"].data_ptr(), s["R"].data_ptr(), (W, H, W), s["ranges"].data_ptr(),
                             s["meta"].data_ptr(), s["st"].data_ptr(), s["u8"].data_ptr(), s["D1"].data_ptr())
    e.stereo_scan_wait()
    sp = scan_lib.ScanPort()
    A = cal.arrays()
    gate = sp.gate(A["Q"], A["XR"], A["XT"], W, H)
    for call, s in enumerate(sets):
        st = s["st"].numpy()
        for f in range(n):
            I1, I2 = s["L"].numpy()[f], s["R"].numpy()[f]
            if call == 2 and f == 1:
                assert st[f] == 1 and (s["D1"].numpy()[f] == 0.0).all()       # zero map, as the reference's caller sees it
                continue
            R1, _ = oracle.process(ol.robotics(dm), I1, I2)
            assert st[f] == 0 and np.array_equal(s["D1"].numpy()[f], R1), (call, f)
            u8_ref = sp.convert_u8(R1)
            assert np.array_equal(s["u8"].numpy()[f], u8_ref)
            r_ref, m_ref = sp.scan(A["Q"], A["XR"], A["XT"], gate, u8_ref)
            r = s["ranges"].numpy()[f]
            assert np.array_equal(r < 1e9 - 1, r_ref < 1e9 - 1) and np.allclose(r, r_ref, rtol=0, atol=1e-9)
    e.close(); sc.close()


def test_rolling_device_submissions_equal_stream_api(jn, synth):
    """jn_stereo_scan_submit_device: four submissions of five device-resident frames rolling through the
    library's sub-batch streams (odd batch: sub-batches of 3 and 2 frames, one textureless frame) give byte
    for byte what jn_elas_process_batch + jn_scan_from_disparity_batch give on a caller stream."""
    import torch
    W, H, dm, n = 320, 240, 64, 5
    dev = torch.device("cuda", 0)
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(scan_lib.fixtures()["Q"]["320x180"])
    sc = jn.ObstacleScan(cal, W, H)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    ins, outs = [], []
    for call in range(4):
        L, R = synth.scene_batch("textured" if call & 1 else "random_dot", W, H, dm, [200 + 7 * call + i for i in range(n)])
        if call == 2:
            L[3] = 9; R[3] = 9
        ins.append((torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)))
        outs.append(dict(D1=torch.full((n, H, W), 5.0, device=dev), st=torch.full((n,), -9, dtype=torch.int32, device=dev),
                         ranges=torch.zeros((n, 90), dtype=torch.float64, device=dev),
                         meta=torch.zeros((n, 5), dtype=torch.float64, device=dev),
                         u8=torch.zeros((n, H, W), dtype=torch.uint8, device=dev)))
    torch.cuda.synchronize()
    for (dL, dR), o in zip(ins, outs):
        e.stereo_scan_submit_device(sc, n, dL.data_ptr(), dR.data_ptr(), (W, H, W), o["D1"].data_ptr(), o["st"].data_ptr(),
                                    o["ranges"].data_ptr(), o["meta"].data_ptr(), o["u8"].data_ptr())
    e.stereo_scan_wait()
    got = [{k: v.cpu().numpy().copy() for k, v in o.items()} for o in outs]
    # the same frames through the stream API (fresh handle: nothing shared with the rolling lanes)
    e2 = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    for call, (dL, dR) in enumerate(ins):
        D1 = torch.zeros((n, H, W), device=dev); st = torch.zeros((n,), dtype=torch.int32, device=dev)
        rg = torch.zeros((n, 90), dtype=torch.float64, device=dev); mt = torch.zeros((n, 5), dtype=torch.float64, device=dev)
        u8 = torch.zeros((n, H, W), dtype=torch.uint8, device=dev)
        e2.process_batch(dL.data_ptr(), dR.data_ptr(), D1.data_ptr(), 0, st.data_ptr(), (W, H, W), n, 0)
        torch.cuda.synchronize()
        stn = st.cpu().numpy()
        D1[torch.from_numpy(stn != 0).to(dev)] = 0.0            # the zero map the submit calls hand out for an unmatched frame
        sc.from_disparity_batch(n, D1.data_ptr(), rg.data_ptr(), mt.data_ptr(), u8.data_ptr(), 0)
        torch.cuda.synchronize()
        g = got[call]
        assert np.array_equal(g["st"], stn) and (stn[3] == 1) == (call == 2), call
        assert np.array_equal(g["D1"], D1.cpu().numpy()), call
        assert np.array_equal(g["u8"], u8.cpu().numpy()), call
        assert np.array_equal(g["ranges"].view(np.int64), rg.cpu().numpy().view(np.int64)), call
        assert np.array_equal(g["meta"].view(np.int64), mt.cpu().numpy().view(np.int64)), call
    e.close(); e2.close(); sc.close()


def test_forty_calls_on_one_handle_leave_no_state_behind(jn, oracle, synth):
    """Workspace buffers are reused from call to call (plane maps, labels, candidate images, grids): 40 calls on one
    handle, scenes alternating so that triangles and uncovered areas move, every result bit-exact."""
    W, H, dm = 200, 150, 48
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    D1 = np.zeros((H, W), np.float32)
    refs = {}
    for call in range(40):
        seed, scene = 300 + call % 5, ("textured" if call % 2 else "random_dot")
        L, R = synth.scene_batch(scene, W, H, dm, [seed])
        I1, I2 = L[0].copy(), R[0].copy()
        if call % 3 == 0:                       # a textureless band: pixels no triangle of THIS call covers
            I1[: H // 3] = 128; I2[: H // 3] = 128
        key = (seed, scene, call % 3 == 0)
        if key not in refs:
            refs[key] = oracle.process(ol.robotics(dm), I1, I2)[0]
        D1[:] = 7
        rc = e.process(I1, I2, D1, None, (W, H, W))
        assert rc == 0 and np.array_equal(D1, refs[key]), call
    e.close()


def test_c4_seeds_full_size(jn, oracle, synth):
    """BASELINE config C4: 1920x1200 pairs with seeds 1000.. -- 32 of them (spread over the 1024),
    final maps bit-exact against the oracle, in one batch through the device-resident entry point."""
    import torch
    W, H, dm = 1920, 1200, 255
    seeds = [1000 + 33 * i for i in range(32)]
    L, R = synth.synth_batch(W, H, dm, seeds)
    dev = torch.device("cuda", 0)
    dL = torch.from_numpy(L).to(dev); dR = torch.from_numpy(R).to(dev)
    dD1 = torch.zeros((len(seeds), H, W), dtype=torch.float32, device=dev)
    st = torch.zeros(len(seeds), dtype=torch.int32, device=dev)
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    e.process_batch(dL.data_ptr(), dR.data_ptr(), dD1.data_ptr(), 0, st.data_ptr(), (W, H, W), len(seeds), 0)
    torch.cuda.synchronize()
    D1 = dD1.cpu().numpy()
    assert (st.cpu().numpy() == 0).all()
    p = ol.robotics(dm)
    for i in range(len(seeds)):
        R1, _ = oracle.process(p, L[i], R[i])
        assert np.array_equal(D1[i], R1), seeds[i]
    e.close()


def test_jpeg_decode_matches_imdecode(jn):
    """nvJPEG luma plane against cv2.imdecode(IMREAD_GRAYSCALE) of the same bitstreams (grayscale and colour
    JPEGs, two qualities): decoders round the IDCT differently, so within 2 grey levels, mean below 0.3."""
    import os
    import torch
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "jpeg_cv2.npz"))
    n = int(z["n"])
    dec = jn.JpegDecoder()
    streams = [z["jpg%d" % i].tobytes() for i in range(n)]
    assert dec.info(streams[0]) == (320, 240)
    dst = torch.zeros((n, 240, 320), dtype=torch.uint8, device="cuda")
    dec.decode_gray_batch(streams, dst.data_ptr(), 320, 240)
    torch.cuda.synchronize()
    got = dst.cpu().numpy().astype(np.int32)
    for i in range(n):
        d = np.abs(got[i] - z["gray%d" % i].astype(np.int32))
        assert d.max() <= 2 and d.mean() < 0.3, (i, d.max(), d.mean())
    # second batch of another size through the same decoder, into a strided destination
    dst2 = torch.zeros((2, 240, 384), dtype=torch.uint8, device="cuda")
    dec.decode_gray_batch(streams[:2], dst2.data_ptr(), 320, 240, dst_stride=384)
    torch.cuda.synchronize()
    assert np.array_equal(dst2.cpu().numpy()[:, :, :320], got[:2].astype(np.uint8))
    with pytest.raises(jn.JnError):
        dec.decode_gray_batch([b"not a jpeg"], dst.data_ptr(), 320, 240)
    dec.close()


def test_pointcloud_batch_equals_single_frame_path(jn, oracle, synth):
    """jn_pointcloud_batch (device resident, batched -g path) against the single-frame host entry point,
    which is checked against the restatement of point_cloud.cpp:298-404 / 149-211 elsewhere: same points in
    the same order (Point32 bits), same packed rgb, same scan."""
    import torch
    W, H, dm, n = 320, 240, 64, 3
    cal = jn.Calibration(scan_lib.CALIB_YML)
    cal.set_q_matrix(scan_lib.fixtures()["Q"]["320x180"])
    sc = jn.ObstacleScan(cal, W, H)
    rng = np.random.default_rng(8)
    Ds, imgs = [], []
    for f in range(n):
        I1, I2, _ = synth.synth_pair(W, H, dm, 90 + f)
        D1, _ = oracle.process(ol.robotics(dm), I1, I2)
        Ds.append(D1); imgs.append(rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
    Ds[1][:] = -10.0                                            # an empty frame in the middle
    dD = torch.from_numpy(np.stack(Ds)).cuda()
    dI = torch.from_numpy(np.stack(imgs)).cuda()
    xyz = torch.zeros((n, W * H, 3), dtype=torch.float32, device="cuda")
    rgb = torch.zeros((n, W * H), dtype=torch.float32, device="cuda")
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    rg = torch.zeros((n, 90), dtype=torch.float64, device="cuda")
    mt = torch.zeros((n, 5), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    for rep in range(2):                                        # second call: nothing left to allocate
        sc.pointcloud_batch(n, dD.data_ptr(), xyz.data_ptr(), cnt.data_ptr(), rg.data_ptr(), mt.data_ptr(), rgb.data_ptr(),
                            dI.data_ptr(), 3 * W, 3, st.cuda_stream)
    torch.cuda.synchronize()
    counts = cnt.cpu().numpy()
    assert counts[1] == 0
    for f in range(n):
        x_ref, c_ref, r_ref, m_ref = sc.pointcloud(Ds[f], imgs[f])
        assert counts[f] == len(x_ref)
        assert np.array_equal(xyz[f, :counts[f]].cpu().numpy().view(np.int32), x_ref.view(np.int32))
        assert np.array_equal(rgb[f, :counts[f]].cpu().numpy().view(np.int32), c_ref.view(np.int32))
        assert np.array_equal(rg[f].cpu().numpy(), r_ref)
    sc.close()
