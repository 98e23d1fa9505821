#!/usr/bin/env python
"""Times one 1920x1200 frame of the reference ELAS on one core of THIS machine at two sets of compiler
flags: -O3 -msse3 (oracle/_ref, the bench baseline) and the reference's literal -msse3 without -O
(CMakeLists.txt:199; `make -C oracle ref_literal`).  SURVEY 8(d): report both.  CPU only.

    make -C oracle ref ref_literal && python tools/time_ref_flags.py
"""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol


def main():
    synth = importlib.import_module("jackal-navigation_b200.synth")
    W, H, dm = 1920, 1200, 255
    I1, I2, _ = synth.synth_pair(W, H, dm, 1000)
    p = ol.robotics(dm)
    maps = {}
    for name, path in (("-O3 -msse3", ol.REF_SO),
                       ("-msse3 (literal)", os.path.join(ROOT, "oracle", "_ref", "literal", "libelas_ref.so"))):
        o = ol.Oracle("ref")
        o.lib = C.CDLL(path)
        o.lib.ref_set_deterministic_heap(0)
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            D1, _ = o.process(p, I1, I2)
            ts.append(time.perf_counter() - t0)
        maps[name] = D1
        print("%-18s median of 3 after a warm-up: %.3f s per frame (%.2f frames/s per core)" % (
            name, sorted(ts[1:])[1], 1.0 / sorted(ts[1:])[1]))
    a, b = maps.values()
    print("maps equal across the two builds:", bool(np.array_equal(a, b)))


if __name__ == "__main__":
    main()
