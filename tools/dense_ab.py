#!/usr/bin/env python
"""A/B timing of kernel variants (tools/build_variant.sh) on the GPU box, cheap enough for a one-minute call.

    python tools/dense_ab.py [--frames 8] [--scene random_dot] LIB [LIB ...]

Synthesises a few 1920x1200 pairs once, then for every library (a path, or "default") runs a child process that
pushes a batch of 32 frames through jn_elas_process_batch with stage profiling on and prints the per-stage event
times (ms per 64 frames) and a digest of the D1 maps: equal digests = bit-equal outputs across the variants."""
import argparse
import ctypes as C
import hashlib
import importlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = ["descriptor", "support", "delaunay", "planes_grid", "raster", "dense_match", "post"]


def child(npz, cfg):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    jn = importlib.import_module("jackal-navigation_b200")
    z = np.load(npz)
    L, R = z["L"], z["R"]
    B = 32
    reps = B // L.shape[0]
    dL = torch.from_numpy(np.concatenate([L] * reps)).cuda()
    dR = torch.from_numpy(np.concatenate([R] * reps)).cuda()
    H, W = L.shape[1:]
    kw = {"filter_median": 1, "postprocess_only_left": 0} if cfg == "c5" else {}
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=255, **kw), device=0)
    dD = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    dS = torch.zeros(B, dtype=torch.int32, device="cuda")
    lib = jn.lib()
    lib.jn_elas_profile.argtypes = [C.c_void_p, C.c_int]
    lib.jn_elas_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    dims = (W, H, W)
    for _ in range(3):
        e.process_batch(dL.data_ptr(), dR.data_ptr(), dD.data_ptr(), 0, dS.data_ptr(), dims, B, 0)
    torch.cuda.synchronize()
    lib.jn_elas_profile(e._h, 1)
    acc = np.zeros(7)
    n = 6
    for _ in range(n):
        e.process_batch(dL.data_ptr(), dR.data_ptr(), dD.data_ptr(), 0, dS.data_ptr(), dims, B, 0)
        torch.cuda.synchronize()
        buf = (C.c_float * 7)()
        lib.jn_elas_profile_read(e._h, buf)
        acc += np.array(list(buf))
    lib.jn_elas_profile(e._h, 0)
    d = hashlib.sha1(dD.cpu().numpy().tobytes()).hexdigest()[:16]
    ok = int((dS.cpu().numpy() == 0).sum())
    ms = acc / n * (64.0 / B)
    print("%-28s %s  D1 %s  ok %d/%d" % (os.path.basename(os.environ.get("JN_ELAS_LIB", "default")),
                                         "  ".join("%s %.3f" % (k[:5], v) for k, v in zip(STAGES, ms)), d, ok, B), flush=True)
    e.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--scene", default="random_dot")
    ap.add_argument("--config", default="robotics")
    ap.add_argument("--child", default=None)
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        return child(a.child, a.config)
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    L, R = bench.make_frames(a.scene, 1920, 1200, 255, [1000 + i for i in range(a.frames)], os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as td:
        npz = os.path.join(td, "frames.npz")
        np.savez(npz, L=L, R=R)
        for lib in a.libs or ["default"]:
            env = dict(os.environ)
            if lib != "default":
                env["JN_ELAS_LIB"] = os.path.abspath(lib)
            else:
                env.pop("JN_ELAS_LIB", None)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", npz, "--config", a.config], env=env)


if __name__ == "__main__":
    main()
