"""World-size-2 gloo test of the frame sharding + final gather (the only N>1 logic of this
path: frames are independent, there is no data-path collective).  CPU only: the per-frame
work is done by the oracle port here; on GPUs bench.py does the same with the CUDA path."""
import importlib
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as ol
    import scan_lib
    sharding = importlib.import_module("jackal-navigation_b200.sharding")
    synth = importlib.import_module("jackal-navigation_b200.synth")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.frame_shard(n_frames, rank, world)
    o = ol.load("port")
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    Q = np.array(fx["Q"]["320x180"]); XR = np.array(fx["calib"]["XR"]); XT = np.array(fx["calib"]["XT"])
    W, H, dm = 160, 120, 32
    gate = sp.gate(Q, XR, XT, W, H)
    rows = []
    for f in range(b, e):
        I1, I2, _ = synth.synth_pair(W, H, dm, 100 + f)
        D1, _ = o.process(ol.robotics(dm), I1, I2)
        r, m = sp.scan(Q, XR, XT, gate, sp.convert_u8(D1))
        rows.append(np.concatenate([r, [m.angle_min, m.angle_max, m.range_min, m.range_max, f]]))
    local = np.array(rows).reshape(-1, 95)
    full = sharding.gather_scans(local, n_frames, rank, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_shard_partitions():
    sharding = importlib.import_module("jackal-navigation_b200.sharding")
    for n in (0, 1, 5, 7, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and b0 <= e0
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_two_ranks_gather_in_frame_order(tmp_path):
    n_frames, world = 5, 2   # ragged: 2 + 3 frames
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert a.shape == (n_frames, 95) and np.array_equal(a, b)
    assert list(a[:, 94]) == [0, 1, 2, 3, 4]          # frame order restored
    # single-process result for the same frames
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    import scan_lib
    synth = importlib.import_module("jackal-navigation_b200.synth")
    o = ol.load("port"); sp = scan_lib.ScanPort(); fx = scan_lib.fixtures()
    Q = np.array(fx["Q"]["320x180"]); XR = np.array(fx["calib"]["XR"]); XT = np.array(fx["calib"]["XT"])
    gate = sp.gate(Q, XR, XT, 160, 120)
    I1, I2, _ = synth.synth_pair(160, 120, 32, 103)
    D1, _ = o.process(ol.robotics(32), I1, I2)
    r, m = sp.scan(Q, XR, XT, gate, sp.convert_u8(D1))
    assert np.array_equal(a[3, :90], r)
