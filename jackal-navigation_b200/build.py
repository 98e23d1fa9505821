"""Builds libjn_elas.so (hand-written CUDA for sm_100a + the C ABI) in-tree.

    python jackal-navigation_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with
the repository snapshot.  Every translation unit is compiled with -fmad=false:
the float/double stages must round like the reference's SSE code (no FMA).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libjn_elas.so")
SOURCES = ["api.cu", "descriptor.cu", "support.cu", "delaunay.cu", "planes_grid.cu", "dense.cu", "post.cu", "scan.cu", "rectify.cu", "navigate.cu", "calib.cu", "jpeg.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return hs


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and not _newer(obj, [path] + _headers()):
        return src, 0, ""
    r = subprocess.run([NVCC] + FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
    return src, r.returncode, (r.stdout + r.stderr)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    logs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, rc, out in ex.map(lambda s: _compile(s, force, verbose), SOURCES):
            logs.append((src, out))
            if rc != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
            if verbose and out:
                print("==", src)
                print(out)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or _newer(LIB, objs):
        # nvJPEG (the decode library behind jn_jpeg_*) is linked statically: nothing to find at run time
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-lnvjpeg_static", "-lculibos"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        for src, out in logs:
            if out:
                f.write("== %s\n%s\n" % (src, out))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
