"""Prints the Delaunay kernel's phase times (globaltimer) for one 1920x1200 frame."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
jn = importlib.import_module("jackal-navigation_b200")
synth = importlib.import_module("jackal-navigation_b200.synth")
W, H, dm = 1920, 1200, 255
I1, I2, _ = synth.synth_pair(W, H, dm, 1000)
e = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm))
D1 = np.zeros((H, W), np.float32); D2 = D1.copy()
for rep in range(3):
    e.process(I1, I2, D1, D2, (W, H, W))
    buf = (C.c_int32 * 8)(); raw = (C.c_char * (32 + 96))()
    jn.lib().jn_elas_frameinfo(C.c_void_p(e._h), 0, raw, 128)
    a = np.frombuffer(raw, np.int32, 8); t = np.frombuffer(raw, np.int64, 12, 32).reshape(2, 6)
    print("n_support %d n_tri %d/%d incon_rounds %d depth %d" % (a[0], a[1], a[2], a[4], a[5]))
    for s in range(2):
        T = t[s]
        print("  side %d: order %.0f us, partition %.0f us, merges+emit %.0f us, total %.0f us" % (
            s, (T[1] - T[0]) / 1e3, (T[2] - T[1]) / 1e3, (T[4] - T[2]) / 1e3, (T[4] - T[0]) / 1e3))
