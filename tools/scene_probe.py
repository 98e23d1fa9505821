"""GPU probe: per-stage device times of a batch for both scene families + Delaunay phase times.

    python tools/scene_probe.py [--batch 16]
"""
import argparse, ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
jn = importlib.import_module("jackal-navigation_b200")
synth = importlib.import_module("jackal-navigation_b200.synth")
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--scenes", default="random_dot,textured")
a = ap.parse_args()
W, H, dm, B = 1920, 1200, 255, a.batch
lib = jn.lib()
lib.jn_elas_profile.argtypes = [C.c_void_p, C.c_int]
lib.jn_elas_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
names = ["descriptor", "support", "delaunay", "planes_grid", "raster", "dense_match", "post"]
for scene in a.scenes.split(","):
    L, R = synth.scene_batch(scene, W, H, dm, [1000 + i for i in range(B)])
    dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
    dD = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    dS = torch.zeros(B, dtype=torch.int32, device="cuda")
    e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
    st = torch.cuda.Stream()
    def step():
        e.process_batch(dL.data_ptr(), dR.data_ptr(), dD.data_ptr(), 0, dS.data_ptr(), (W, H, W), B, st.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    print("%s: %d frames/step, %.3f ms/step, %.1f frames/s" % (scene, B, e0.elapsed_time(e1) / 5, B * 5 / e0.elapsed_time(e1) * 1e3))
    lib.jn_elas_profile(e._h, 1)
    acc = np.zeros(7)
    for _ in range(3):
        step(); torch.cuda.synchronize()
        buf = (C.c_float * 7)(); lib.jn_elas_profile_read(e._h, buf); acc += np.array(list(buf))
    lib.jn_elas_profile(e._h, 0)
    print("   stage ms/step:", ", ".join("%s %.3f" % (n, v / 3) for n, v in zip(names, acc)))
    for f in (0, B - 1):
        raw = (C.c_char * 128)()
        lib.jn_elas_frameinfo(C.c_void_p(e._h), f, raw, 128)
        i = np.frombuffer(raw, np.int32, 8); t = np.frombuffer(raw, np.int64, 12, 32).reshape(2, 6)
        print("   frame %d: n_support %d n_tri %d/%d status %d depth %d" % (f, i[0], i[1], i[2], i[3], i[5]))
        for s in range(2):
            T = t[s]
            print("     side %d: order %.0f us, partition %.0f us, merges+emit %.0f us" % (
                s, (T[1] - T[0]) / 1e3, (T[2] - T[1]) / 1e3, (T[4] - T[2]) / 1e3))
    e.close()
