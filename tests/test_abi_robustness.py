"""The host side of the C ABI takes bad input without crashing: NULL handles and pointers on every entry point
that can be reached without a device return an error code (or do nothing for the void ones), and the calibration
reader (cv::FileStorage's replacement, point_cloud.cpp:530-538) answers mutated files with JN_OK or JN_ERR_IO.
Runs in a child process so that a crash is a test failure and not the end of the test session."""
import os
import subprocess
import sys

import oracle_lib as ol

ROOT = ol.ROOT

CHILD = r'''
import ctypes as C, importlib, random, sys
sys.path.insert(0, %(root)r)
jn = importlib.import_module("jackal-navigation_b200")
l = jn.lib()
P, N = C.c_void_p, None
d = C.c_double
l.jn_calib_load_yaml.argtypes = [C.c_char_p, P]
l.jn_calib_set_q.argtypes = [P, d, d, d, d]
l.jn_calib_stereo_rectify.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, d, P, P, P, P]
l.jn_calib_init_undistort_rectify_map.argtypes = [P, P, P, P, C.c_int, C.c_int, P, P]
l.jn_scan_compact.argtypes = [P, P]
l.jn_scan_create.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
l.jn_elas_params_default.argtypes = [P, C.c_int]
l.jn_elas_create.argtypes = [P, C.c_int]
l.jn_navigate_set_clearance.argtypes = [P, d, d, C.c_int]
l.jn_navigate_set_scan.argtypes = [P, P, C.c_int, d, d]
cal = jn.Calib()
yml = %(yml)r.encode()
ERR = lambda rc: rc is not None and rc < 0
NUL = lambda rc: rc is None
ANY = lambda rc: True
cases = [
    ("destroy(NULL) x4", lambda: (l.jn_elas_destroy(N), l.jn_scan_destroy(N), l.jn_rectify_destroy(N), l.jn_jpeg_destroy(N), l.jn_navigate_destroy(N)), ANY),
    ("params_default(NULL)", lambda: l.jn_elas_params_default(N, 0), ANY),
    ("params_default(bad setting)", lambda: l.jn_elas_params_default(C.byref(jn.parameters()), 7), ANY),
    ("create(NULL params)", lambda: l.jn_elas_create(N, 0), NUL),
    ("process(NULL...)", lambda: l.jn_elas_process(N, N, N, N, N, N), ERR),
    ("process_batch(NULL...)", lambda: l.jn_elas_process_batch(N, 1, N, N, N, N, N, N, N), ERR),
    ("load_yaml(NULL path)", lambda: l.jn_calib_load_yaml(N, C.byref(cal)), ERR),
    ("load_yaml(NULL out)", lambda: l.jn_calib_load_yaml(yml, N), ERR),
    ("load_yaml(missing)", lambda: l.jn_calib_load_yaml(b"/nonexistent/x.yml", C.byref(cal)), ERR),
    ("load_yaml(directory)", lambda: l.jn_calib_load_yaml(b"/tmp", C.byref(cal)), ERR),
    ("set_q(NULL)", lambda: l.jn_calib_set_q(N, 1, 1, 1, 1), ANY),
    ("stereo_rectify(NULL)", lambda: l.jn_calib_stereo_rectify(N, 640, 360, 0, 0, 1, 0.0, N, N, N, N), ERR),
    ("stereo_rectify(0x0)", lambda: l.jn_calib_stereo_rectify(C.byref(cal), 0, 0, 0, 0, 1, 0.0, N, N, N, N), ERR),
    ("stereo_rectify(zeroed calibration)", lambda: l.jn_calib_stereo_rectify(C.byref(jn.Calib()), 640, 360, 0, 0, 1, 0.0, N, N, N, N), ERR),
    ("compose_cam_to_robot(NULL)", lambda: l.jn_calib_compose_cam_to_robot(N, d(0), d(0), d(0), d(0), d(0), d(0)), ERR),
    ("undistort_map(NULL)", lambda: l.jn_calib_init_undistort_rectify_map(N, N, N, N, 4, 4, N, N), ERR),
    ("scan_compact(NULL)", lambda: l.jn_scan_compact(N, N), ERR),
    ("scan_create(NULL calib)", lambda: l.jn_scan_create(N, 64, 48, 0, 0, 0), lambda rc: not rc),
    ("scan_from_disparity(NULL)", lambda: l.jn_scan_from_disparity(N, N, N, N, N), ERR),
    ("scan_from_disparity_batch(NULL)", lambda: l.jn_scan_from_disparity_batch(N, 1, N, N, N, N, N), ERR),
    ("scan_gate_cache(NULL)", lambda: l.jn_scan_gate_cache(N, N), ERR),
    ("points_from_disparity(NULL)", lambda: l.jn_points_from_disparity(N, N, N, N, N, N), ERR),
    ("pointcloud_from_disparity(NULL)", lambda: l.jn_pointcloud_from_disparity(N, N, N, 0, 1, N, N, N, N, N), ERR),
    ("pointcloud_batch(NULL)", lambda: l.jn_pointcloud_batch(N, 1, N, N, 0, 1, N, N, N, N, N, N), ERR),
    ("stereo_scan_submit(NULL)", lambda: l.jn_stereo_scan_submit(N, N, 1, N, N, N, N, N, N, N, N), ERR),
    ("stereo_scan_submit_device(NULL)", lambda: l.jn_stereo_scan_submit_device(N, N, 1, N, N, N, N, N, N, N, N), ERR),
    ("stereo_scan_batch_host(NULL)", lambda: l.jn_stereo_scan_batch_host(N, N, 1, N, N, N, N, N, N, N, N), ERR),
    ("stereo_scan_wait(NULL)", lambda: l.jn_stereo_scan_wait(N), ERR),
    ("rectify_create(NULL maps)", lambda: l.jn_rectify_create(N, N, 4, 4, 0), lambda rc: not rc),
    ("rectify_batch(NULL)", lambda: l.jn_rectify_batch(N, 1, N, 4, 4, 4, N, N, 4, N), ERR),
    ("jpeg_info(NULL)", lambda: l.jn_jpeg_info(N, N, 0, N, N), ERR),
    ("jpeg_decode(NULL)", lambda: l.jn_jpeg_decode_gray_batch(N, 1, N, N, N, 4, 4, 4, 16, N), ERR),
    ("navigate_*(NULL)", lambda: (l.jn_navigate_set_clearance(N, 1.0, 1.0, 1), l.jn_navigate_set_last_dir(N, 1), l.jn_navigate_last_dir(N)), ANY),
    ("navigate_set_scan(NULL)", lambda: l.jn_navigate_set_scan(N, N, 0, 0.0, 0.0), ERR),
    ("navigate_set_scan_bins(NULL)", lambda: l.jn_navigate_set_scan_bins(N, N, N), ERR),
    ("navigate_check_obstacle(NULL)", lambda: l.jn_navigate_check_obstacle(N, N), ERR),
    ("navigate_command(NULL)", lambda: l.jn_navigate_command(N, 1, d(0), d(0), N), ERR),
    ("navigate_points(NULL)", lambda: l.jn_navigate_points(N, N, 0), ERR),
    ("navigate_choose_direction(NULL)", lambda: l.jn_navigate_choose_direction(N), ERR),
    ("host_free(NULL)", lambda: l.jn_host_free(N), ANY),
    ("cache_clear()", lambda: l.jn_cache_clear(), ANY),
    ("profile(NULL)", lambda: l.jn_elas_profile(N, 1), ERR),
    ("profile_read(NULL)", lambda: l.jn_elas_profile_read(N, N), ERR),
]
for name, f, ok in cases:
    print("CASE", name, flush=True)
    rc = f()
    assert ok(rc), (name, rc)
print("NULL-OK", len(cases), flush=True)

src = open(yml, "rb").read()
random.seed(20261017)
path = sys.argv[1]
seen = {}
for it in range(400):
    b = bytearray(src)
    mode = it %% 6
    if mode == 0:
        b = b[:random.randrange(len(b))]
    elif mode == 1:
        for _ in range(random.randrange(1, 20)):
            b[random.randrange(len(b))] = random.randrange(256)
    elif mode == 2:
        i = random.randrange(len(b)); del b[i:i + random.randrange(1, 200)]
    elif mode == 3:
        i = random.randrange(len(b)); b[i:i] = bytes(random.choice(b"[],:-0123456789.e \n") for _ in range(random.randrange(1, 300)))
    elif mode == 4:
        b = b.replace(b"rows: 3", b"rows: %%d" %% random.choice([0, -1, 1, 99999999, 2 ** 31 - 1]), random.randrange(1, 4))
    else:
        b = b.replace(b"data:", random.choice([b"data", b"dat:", b"data: [", b"data: ]"]), random.randrange(1, 4))
    open(path, "wb").write(b)
    print("YAML", it, flush=True)
    rc = l.jn_calib_load_yaml(path.encode(), C.byref(jn.Calib()))
    assert rc in (0, -4), rc
    seen[rc] = seen.get(rc, 0) + 1
assert seen.get(-4, 0) > 100 and seen.get(0, 0) > 0, seen
print("YAML-OK", seen, flush=True)
'''


def test_host_entry_points_survive_null_and_mutated_calibration_files(jn, tmp_path):
    code = CHILD % {"root": ROOT, "yml": os.path.join(ROOT, "tests", "golden", "calib_c920.yml")}
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path / "fuzz.yml")], capture_output=True, text=True,
                       timeout=300)
    last = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
    assert r.returncode == 0, "child died (rc %d) at: %s\n%s" % (r.returncode, last, r.stderr[-1500:])
    assert "NULL-OK" in r.stdout and "YAML-OK" in r.stdout
