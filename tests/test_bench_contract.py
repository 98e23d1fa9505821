"""bench.py's JSON contract, checked without a GPU: tools/bench_dryrun.py executes the GPU arm's host-side flow
with torch.cuda and the library handles replaced by stand-ins (no numbers are measured), the reference arm runs
for real on a small frame.  Each run is a subprocess: the dry run patches torch."""
import json
import os
import subprocess
import sys

import oracle_lib as ol

ROOT = ol.ROOT
SMALL = ["--width", "320", "--height", "240", "--disp-max", "64", "--batch", "4", "--steps", "2", "--warmup", "1"]
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"]


def _json_line(cmd, env=None):
    r = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_gpu_arm_prints_every_contract_key():
    d = _json_line(["tools/bench_dryrun.py"] + SMALL)
    for k in BASE_KEYS + ["gpu_launches", "roofline", "clocks", "cpu_baseline", "parity", "stage_ms_per_step",
                          "stage_roofline", "single_frame_latency_ms", "scenes", "c5"]:
        assert k in d, k
    assert d["metric"] == "elas_stereo_to_obstacle_scan_throughput" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3
    assert "model" not in d["config"] and d["config"]["workload"].startswith("320x240 disp_max=64 ROBOTICS")
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 4 * 320 * 240          # both images of every frame of the step
    assert d["e2e"]["d2h_bytes_per_step"] == 4 * (90 * 8 + 40 + 4)      # scans + meta + status
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the stand-ins compute nothing, so the in-bench checker must say the maps differ: it really compares
    assert d["parity"]["frames_checked"] == 4 and d["parity"]["ok"] is False and d["parity"]["d1_maps_bit_equal"] == 0


def test_gpu_arm_strong_scaling_flow():
    """BASELINE config C4: --total-frames T split contiguously, one pass over the shard, extras off."""
    d = _json_line(["tools/bench_dryrun.py"] + SMALL + ["--total-frames", "12", "--no-cpu-baseline"])
    assert d["scaling"] == "strong" and d["config"]["total_frames"] == 12
    assert d["steps"] == 3                                  # 12 frames in batches of 4
    assert "scenes" not in d and "c5" not in d
    assert abs(d["value"] - 12 / (15.0 / 1000.0)) < 1e-6    # the stand-in events report 15 ms


def test_gpu_arm_multi_rank_flow():
    """One process as rank 0 and as rank 1 of a two-rank launch (collectives replaced by no-ops): rank 0 prints the
    line with the whole-job value and without the single-GPU sections, rank 1 prints nothing."""
    env = {"WORLD_SIZE": "2", "LOCAL_RANK": "0", "RANK": "0"}
    d = _json_line(["tools/bench_dryrun.py"] + SMALL + ["--gpus", "2"], env)
    assert d["n_gpus"] == 2 and d["scaling"] == "weak"
    assert abs(d["value"] - 2 * 4 * 2 / (15.0 / 1000.0)) < 1e-6          # 2 ranks x 4 frames x 2 steps in 15 ms
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-6
    for k in ("scenes", "c5", "single_frame", "cpu_baseline", "parity"):
        assert k not in d, k                                             # sections of the single-GPU line only
    assert "2 GPU(s)" in d["config"]["parallelism"] and d["e2e"]["rank_pinning"] is not None
    r = subprocess.run([sys.executable, "tools/bench_dryrun.py"] + SMALL + ["--gpus", "2"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, WORLD_SIZE="2", LOCAL_RANK="1", RANK="1"))
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")], r.stderr[-1500:]


def test_reference_arm_prints_the_same_config_and_the_reference_keys():
    ours = _json_line(["tools/bench_dryrun.py"] + SMALL + ["--no-cpu-baseline", "--no-extras"])
    ref = _json_line(["bench.py", "--impl", "reference"] + SMALL)
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in ref, k
    assert ref["impl"] == "reference" and ref["config"] == ours["config"]
    for k in ("metric", "unit", "higher_is_better", "scaling", "dtype"):
        assert ref[k] == ours[k], k
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert ref["cpu_baseline"]["value"] == ref["value"] and ref["value"] > 0
    assert ref["steps"] == 2 and ref["warmup"] == 1


def test_reference_arm_other_ranks_exit_without_work():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference"] + SMALL, cwd=ROOT, capture_output=True,
                       text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
