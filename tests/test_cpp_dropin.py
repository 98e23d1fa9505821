"""The C++ drop-in header include/elas.h: a reference-style caller compiles against it
unchanged, links against libjn_elas.so, and (on a GPU) reproduces the oracle's maps."""
import os
import subprocess
import numpy as np
import pytest
import oracle_lib as ol

ROOT = ol.ROOT
SRC = os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")


def build(tmp_path):
    exe = str(tmp_path / "dropin_main")
    libdir = os.path.join(ROOT, "jackal-navigation_b200")
    cmd = ["g++", "-O1", "-std=c++11", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-ljn_elas", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_reference_style_caller_compiles_and_links(jn, tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "ROBOTICS support_threshold=0.85 gamma=3 ipol_gap_width=3" in out
    assert "MIDDLEBURY support_threshold=0.95 gamma=5 ipol_gap_width=5000" in out


def test_cpp_caller_fails_loudly_without_gpu(jn, synth, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = build(tmp_path)
    I1, I2, _ = synth.synth_pair(160, 120, 32, 1)
    I1.tofile(tmp_path / "l.u8"); I2.tofile(tmp_path / "r.u8")
    r = subprocess.run([exe, "160", "120", "32", str(tmp_path / "l.u8"), str(tmp_path / "r.u8"),
                        str(tmp_path / "ol.f32"), str(tmp_path / "or.f32")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cpp_caller_matches_oracle(jn, oracle, synth, tmp_path):
    exe = build(tmp_path)
    W, H, dm = 320, 240, 64
    I1, I2, _ = synth.synth_pair(W, H, dm, 4)
    I1.tofile(tmp_path / "l.u8"); I2.tofile(tmp_path / "r.u8")
    subprocess.run([exe, str(W), str(H), str(dm), str(tmp_path / "l.u8"), str(tmp_path / "r.u8"),
                    str(tmp_path / "ol.f32"), str(tmp_path / "or.f32")], check=True)
    D1 = np.fromfile(tmp_path / "ol.f32", np.float32).reshape(H, W)
    D2 = np.fromfile(tmp_path / "or.f32", np.float32).reshape(H, W)
    R1, R2 = oracle.process(ol.robotics(dm), I1, I2)
    assert np.array_equal(D1, R1) and np.array_equal(D2, R2)


REF_NODE = "/root/reference/src/obstacle_avoidance/point_cloud.cpp"


@pytest.mark.skipif(not os.path.exists(REF_NODE), reason="needs the reference tree")
def test_reference_node_source_compiles_against_the_dropin_header(jn, synth, tmp_path):
    """INTEGRATION.md section 1, literally: the reference's point_cloud.cpp, unmodified and where it lies, compiled
    with include/elas.h ahead of the reference's own elas.h (ROS / OpenCV from oracle/standins) and linked against
    libjn_elas.so instead of the reference's ELAS objects.  generateDisparityMap's `Elas elas(param);
    elas.process(...)` (:416-419) then runs this repository's library; without a GPU it must fail loudly."""
    import sys
    import torch
    so = str(tmp_path / "libpointcloud_dropin.so")
    libdir = os.path.join(ROOT, "jackal-navigation_b200")
    cmd = ["g++", "-O1", "-fPIC", "-w", "-std=gnu++11", "-shared",
           "-I", os.path.join(ROOT, "oracle", "standins"), "-I", os.path.dirname(REF_NODE),
           "-I", os.path.join(ROOT, "include"),                      # our elas.h wins over the reference's
           "-I", "/root/reference/src/elas",                         # image.h
           os.path.join(ROOT, "oracle", "pointcloud_ref_shim.cpp"), "-o", so,
           "-L", libdir, "-ljn_elas", "-Wl,-rpath," + libdir, "-Wl,--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert " U jn_elas_process" in syms and " U jn_elas_create" in syms      # the node calls the product library
    assert "computeSupportMatches" not in syms                               # and carries no ELAS of its own
    if torch.cuda.is_available():
        return
    I1, I2, _ = synth.synth_pair(160, 120, 32, 1)
    I1.tofile(tmp_path / "l.u8"); I2.tofile(tmp_path / "r.u8")
    child = ("import ctypes as C, numpy as np\n"
             "l = C.CDLL(%r)\n"
             "P = C.c_void_p\n"
             "l.ref_pc_setup.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]\n"
             "l.ref_pc_generate_disparity.argtypes = [P, P, P]\n"
             "Q = np.eye(4); XR = np.eye(3); XT = np.zeros(3)\n"
             "p = lambda a: a.ctypes.data_as(P)\n"
             "l.ref_pc_setup(p(Q), p(XR), p(XT), 160, 120, 0, 0)\n"
             "I1 = np.fromfile(%r, np.uint8); I2 = np.fromfile(%r, np.uint8); out = np.zeros(160 * 120, np.uint8)\n"
             "l.ref_pc_generate_disparity(p(I1), p(I2), p(out))\n"
             "print('returned')\n" % (so, str(tmp_path / "l.u8"), str(tmp_path / "r.u8")))
    r = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True)
    assert r.returncode != 0 and "returned" not in r.stdout
    assert "Elas" in r.stderr and "no CPU path" in r.stderr, r.stderr[-500:]


def test_c_abi_is_plain_c99_and_the_example_builds(jn, tmp_path):
    """include/jn_elas.h + jn_elas_debug.h are what an FFI binds: they must be valid strict C99 on their own, and
    examples/camera_to_command.c (calibration -> ELAS -> scan -> vote through host-pointer entry points only) must
    build against them with every warning an error.  Without a GPU the program stops after the calibration step
    with exit code 3 and the library's "no CPU path" message."""
    import torch
    hdr = tmp_path / "hdr.c"
    hdr.write_text('#include "jn_elas.h"\n#include "jn_elas_debug.h"\nint main(void) { return 0; }\n')
    strict = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include")]
    r = subprocess.run(strict + ["-c", str(hdr), "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = str(tmp_path / "camera_to_command")
    libdir = os.path.join(ROOT, "jackal-navigation_b200")
    r = subprocess.run(strict + ["-O1", os.path.join(ROOT, "examples", "camera_to_command.c"), "-o", exe, "-L", libdir,
                                 "-ljn_elas", "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "calib_c920.yml"), "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU path" in r.stderr
    assert r.stdout.startswith("Q: cx ")                       # the OpenCV-free stereoRectify ran on the host
