#!/usr/bin/env python
"""What the STOCK node does where this repository had to define a behaviour.

    python tools/h8_probe.py

Hands the reference's own publishObstacleScan(Mat&) (point_cloud.cpp:213-296, compiled in place as
oracle/_ref/libpointcloud_ref.so) an all-invalid disparity map on a window of the 1920x1200 calibration whose gate
cache wrapped to 0 (cacheDisparityValues stores d = 256 in a uchar, :143).  Disparity 0 passes the wrapped gate,
Q * (i, j, 0, 1) has w = 0, the angle is NaN, (int)floor(NaN) indexes scan[] -- in a child process, because the
observed outcome on this image is a segmentation fault.  The CUDA path and the restatement skip such pixels."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np, scan_lib, ref_nodes_lib as rn
fx = scan_lib.fixtures()
XR, XT, Q = np.array(fx["calib"]["XR"]), np.array(fx["calib"]["XT"]), np.array(fx["Q"]["1920x1200_Kx3"])
W, H, ox, oy = 240, 150, 820, 700
node = rn.PointCloudNode(Q, XR, XT, W, H, ox, oy)
g = node.cache_gate()
u8 = np.zeros((H, W), np.uint8)
n, lo, hi, nan = rn.bin_index_range(Q, XR, XT, u8, gate=g, ox=ox, oy=oy)
print("window %%dx%%d at (%%d, %%d): %%d of %%d gate entries wrapped to 0; %%d pixels pass the gate with d = 0, %%d NaN angles"
      %% (W, H, ox, oy, int((g[..., 0] == 0).sum()), W * H, n, nan), flush=True)
r, m = node.scan(u8)
print("the stock node returned: %%d ranges, meta %%s" %% (len(r), list(m)))
''' % os.path.join(ROOT, "tests")

if __name__ == "__main__":
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True)
    print(r.stdout.strip())
    print("child exit code %d%s" % (r.returncode, " (killed by signal %d)" % -r.returncode if r.returncode < 0 else ""))
