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
