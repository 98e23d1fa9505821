#!/usr/bin/env python
"""bench.py -- stereo-to-obstacle-scan throughput at 1920x1200, disp_max 255 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one pass of the whole hot path (Elas::process pipeline + fused
convert/reproject/transform/scan) over one batch of B synthetic random-dot stereo pairs.
One process per GPU; for N > 1 launch under torchrun (the frames are independent, so ranks
share nothing: weak scaling, no collective on the data path -- NCCL is used for the barrier
and the max-over-ranks of the device time only).

Prints ONE JSON line (rank 0):
  value            frames/s over all ranks, inputs resident in HBM, CUDA-event timed
  e2e              same metric through the C ABI with HOST (pinned) buffers: H2D of the
                   image pairs and D2H of the scan + u8 disparity map inside the timed region
  roofline         dense-matching kernel: algorithmic bytes (72*W*H per frame) / its event time
  cpu_baseline     the reference's own ELAS (oracle/_ref, built from /root/reference) on the
                   host cores, one frame per core in separate processes
--impl reference times that CPU arm alone, same metric / config.
"""
import argparse
import ctypes as C
import importlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "elas_stereo_to_obstacle_scan_throughput"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--disp-max", type=int, default=255)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames-per-core", type=int, default=2)
    return ap.parse_args()


def workload_name(a):
    return "%dx%d disp_max=%d ROBOTICS(postprocess_only_left) ELAS + C920xK3 reproject/XR,XT/90-bin scan" % (
        a.width, a.height, a.disp_max)


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process = one core: runs the reference ELAS on `n` frames (Triangle keeps mutable
    file-scope state, so cores are separate processes)."""
    W, H, dm, seeds, kind = args
    import numpy as np
    import oracle_lib as ol
    synth = importlib.import_module("jackal-navigation_b200.synth")
    import scan_lib
    o = ol.load(kind)
    if kind == "ref":
        o.lib.ref_set_deterministic_heap(0)   # timing run: default allocator
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    Q = np.array(fx["Q"]["1920x1200_Kx3"] if (W, H) == (1920, 1200) else fx["Q"]["640x480"])
    XR = np.array(fx["calib"]["XR"]); XT = np.array(fx["calib"]["XT"])
    frames = [synth.synth_pair(W, H, dm, s)[:2] for s in seeds]
    gate = np.zeros((H, W, 2), np.uint8)
    gate[..., 0] = 3
    gate[..., 1] = 255   # init-time cache is not part of the per-frame path
    p = ol.robotics(dm)
    t0 = time.perf_counter()
    for I1, I2 in frames:
        D1, _ = o.process(p, I1, I2)
        sp.scan(Q, XR, XT, gate, sp.convert_u8(D1))
    return time.perf_counter() - t0, len(frames)


def cpu_arm(a, frames_per_core, cores=None):
    import oracle_lib as ol
    kind = "ref" if os.path.exists(ol.REF_SO) else "port"
    cores = cores or (os.cpu_count() or 1)
    # single frame on one core
    t1, n1 = _cpu_worker((a.width, a.height, a.disp_max, [5000], kind))
    jobs = [(a.width, a.height, a.disp_max, [6000 + c * 16 + k for k in range(frames_per_core)], kind)
            for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    nfr = sum(r[1] for r in res)
    return {"value": nfr / busy, "unit": UNIT, "cores": cores,
            "kind": "reference" if kind == "ref" else "port",
            "sample": "%d frames (%d per core, one process per core), max worker time %.2fs, pool wall %.2fs; "
                      "single frame on one core: %.3fs (%.3f frames/s); ELAS built -O3 -msse3" % (
                          nfr, frames_per_core, busy, wall, t1 / n1, n1 / t1),
            "single_core_frames_per_s": n1 / t1}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cb = None
    for i in range(a.warmup + a.steps):
        cb = cpu_arm(a, 1)
        if i >= a.warmup:
            vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * cb["cores"] / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(a), "step": "one frame per host core (%d cores)" % cb["cores"]},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    jn = importlib.import_module("jackal-navigation_b200")
    synth = importlib.import_module("jackal-navigation_b200.synth")
    import scan_lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, dm, B = a.width, a.height, a.disp_max, a.batch
    n = W * H

    # ---- synthetic inputs: B distinct pairs per rank (seeds as SURVEY C4: 1000 + global index)
    L, R = synth.synth_batch(W, H, dm, [1000 + rank * B + i for i in range(B)])
    hL = torch.from_numpy(L).pin_memory()
    hR = torch.from_numpy(R).pin_memory()
    dL = hL.to(dev); dR = hR.to(dev)
    dD1 = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    dStatus = torch.zeros(B, dtype=torch.int32, device=dev)
    dRanges = torch.empty((B, 90), dtype=torch.float64, device=dev)
    dMeta = torch.empty((B, 5), dtype=torch.float64, device=dev)   # jn_scan_meta = 40 bytes
    dU8 = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    hRanges = torch.empty((B, 90), dtype=torch.float64).pin_memory()
    hMeta = torch.empty((B, 5), dtype=torch.float64).pin_memory()
    hU8 = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()

    elas = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm), device=local)
    cal = jn.Calibration(scan_lib.CALIB_YML)
    fx = scan_lib.fixtures()
    cal.set_q_matrix(fx["Q"]["1920x1200_Kx3"] if (W, H) == (1920, 1200) else fx["Q"]["640x480"])
    scan = jn.ObstacleScan(cal, W, H, device=local)
    dims = (W, H, W)
    stream = torch.cuda.Stream(device=dev)
    copy_stream = torch.cuda.Stream(device=dev)

    def step_resident():
        elas.process_batch(dL.data_ptr(), dR.data_ptr(), dD1.data_ptr(), 0, dStatus.data_ptr(), dims, B,
                           stream.cuda_stream)
        scan.from_disparity_batch(B, dD1.data_ptr(), dRanges.data_ptr(), dMeta.data_ptr(), dU8.data_ptr(),
                                  stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            if finish:
                finish()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput
    with torch.cuda.stream(stream):
        for _ in range(max(a.warmup, 3)):
            step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = jn.launch_count()
    ms = timed(step_resident, a.steps)
    launches = jn.launch_count() - l0
    status = dStatus.cpu().numpy()
    value = world * B * a.steps / (ms / 1000.0)

    # ---- end to end with host buffers: H2D + pipeline + D2H every step (double-buffered inputs)
    dL2 = [torch.empty_like(dL) for _ in range(2)]
    dR2 = [torch.empty_like(dR) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    it = [0]

    out_stream = torch.cuda.Stream(device=dev)
    dRanges2 = [torch.empty_like(dRanges) for _ in range(2)]
    dMeta2 = [torch.empty_like(dMeta) for _ in range(2)]
    dU82 = [torch.empty_like(dU8) for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]

    def step_e2e():
        # three streams: H2D of step i+1, compute of step i and D2H of step i-1 overlap;
        # inputs and outputs are double-buffered on the device
        k = it[0] & 1
        it[0] += 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[k])        # input buffer k free again
            dL2[k].copy_(hL, non_blocking=True)
            dR2[k].copy_(hR, non_blocking=True)
            ev_in[k].record(copy_stream)
        stream.wait_event(ev_in[k])
        stream.wait_event(ev_copied[k])               # output buffer k already read back
        elas.process_batch(dL2[k].data_ptr(), dR2[k].data_ptr(), dD1.data_ptr(), 0, dStatus.data_ptr(), dims, B,
                           stream.cuda_stream)
        scan.from_disparity_batch(B, dD1.data_ptr(), dRanges2[k].data_ptr(), dMeta2[k].data_ptr(),
                                  dU82[k].data_ptr(), stream.cuda_stream)
        ev_done[k].record(stream)
        ev_out[k].record(stream)
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(ev_out[k])
            hRanges.copy_(dRanges2[k], non_blocking=True)
            hMeta.copy_(dMeta2[k], non_blocking=True)
            hU8.copy_(dU82[k], non_blocking=True)
            ev_copied[k].record(out_stream)

    def drain_e2e():
        for k in range(2):                            # the timed region ends when the last D2H has landed
            stream.wait_event(ev_copied[k])

    for k in range(2):
        ev_done[k].record(stream)
        ev_copied[k].record(stream)
    with torch.cuda.stream(stream):
        for _ in range(3):
            step_e2e()
        drain_e2e()
    torch.cuda.synchronize()
    ms_e2e = timed(step_e2e, a.steps, drain_e2e)
    clocks = sampler.stop() if sampler else None     # sampled over both timed regions
    e2e = world * B * a.steps / (ms_e2e / 1000.0)
    h2d = 2 * B * n
    d2h = B * (n + 90 * 8 + 40)

    # ---- per-stage device times (CUDA events on the launching stream) for the roofline
    lib = jn.lib()
    lib.jn_elas_profile.argtypes = [C.c_void_p, C.c_int]
    lib.jn_elas_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.jn_elas_profile(elas._h, 1)
    stages = np.zeros(7, np.float64)
    reps = 3
    for _ in range(reps):
        with torch.cuda.stream(stream):
            step_resident()
        torch.cuda.synchronize()
        buf = (C.c_float * 7)()
        lib.jn_elas_profile_read(elas._h, buf)
        stages += np.array(list(buf))
    stages /= reps
    lib.jn_elas_profile(elas._h, 0)
    names = ["descriptor", "support", "delaunay", "planes_grid", "raster", "dense_match", "post"]

    # ---- latency of the reference-facing single-frame call (host numpy buffers, synchronous)
    lat = None
    if rank == 0:
        e1 = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=dm), device=local)
        D1h = np.zeros((H, W), np.float32); D2h = np.zeros((H, W), np.float32)
        ts = []
        for i in range(8):
            t0 = time.perf_counter()
            e1.process(L[0], R[0], D1h, D2h, dims)
            ts.append(time.perf_counter() - t0)
        lat = 1000.0 * sorted(ts[2:])[len(ts[2:]) // 2]
        e1.close()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dense_ms = float(stages[5])
    achieved = 72.0 * n * B / (dense_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dense_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_frame") * B   # per launch of B frames, like `achieved`
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step_per_gpu": B, "pairs": "distinct random-dot pairs, "
                   "seeds 1000+", "l2": "inputs per step (%.0f MB) and working set (>4 GB) exceed the 126 MB L2" % (
                       2 * B * n / 1e6), "parallelism": "frames sharded across %d GPU(s), no data-path collective" % world},
        "mpix_per_s": value * n / 1e6, "ms_per_frame": ms / a.steps / B,
        "frames_ok": int((status == 0).sum()), "frames_few_support": int((status == 1).sum()),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / a.steps,
                "what": "pinned host image pairs -> H2D -> jn_elas_process_batch + jn_scan_from_disparity_batch -> "
                        "D2H of 90-bin scans, scan meta and the u8 disparity maps (copies on their own streams, "
                        "double-buffered, all inside the timed region)"},
        "gpu_launches": int(launches),
        "single_frame_latency_ms": lat,   # jn_elas_process: H2D + pipeline + D2H of both maps, one frame
        "stage_ms_per_step": {k: float(v) for k, v in zip(names, stages)},
        "roofline": {"bound": "hbm", "kernel": "dense_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "algorithmic_bytes_per_launch": 72.0 * n * B, "kernel_ms": dense_ms},
        "clocks": clocks,
    }
    if not a.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_arm(a, a.cpu_frames_per_core)
        except Exception as ex:   # the checker is optional for the GPU arm
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
