#!/usr/bin/env python
"""bench.py -- stereo-to-obstacle-scan throughput at 1920x1200, disp_max 255 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                    [--scene random_dot|textured] [--config robotics|c5] [--total-frames T]

A "step" is one pass of the whole hot path (Elas::process pipeline + fused
convert/reproject/transform/scan) over one batch of B synthetic stereo pairs.
One process per GPU; for N > 1 launch under torchrun (the frames are independent, so ranks
share nothing: no collective on the data path -- NCCL is used for the barrier and the
max-over-ranks of the device time only).  Default: B distinct pairs per GPU per step ("weak").
--total-frames T: BASELINE config C4 as written -- T pairs (seeds 1000..1000+T-1) split contiguously
over the ranks, every rank works through its shard in batches of B ("strong").

Prints ONE JSON line (rank 0):
  value            frames/s over all ranks, inputs resident in HBM, CUDA-event timed
  e2e              same metric through the C ABI's host-buffer entry point
                   (jn_stereo_scan_submit / _wait): pinned host image pairs in, scans + u8 maps out,
                   H2D and D2H inside the timed region
  roofline         dense-matching kernel: algorithmic bytes (72*W*H per frame) / its event time
  parity           frames of the timed batch re-computed by the CPU checker: maps and scans equal
  scenes, c5       the same measurement on the second scene family / BASELINE config C5 (short runs; N = 1 only)
  cpu_baseline     the reference's own ELAS (oracle/_ref, built from /root/reference) on the
                   host cores, one frame per core in separate processes
--impl reference times that CPU arm alone, same metric / config.
"""
import argparse
import ctypes as C
import hashlib
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
SEED0 = 1000          # seeds 1000 + global frame index


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
    ap.add_argument("--scene", default="random_dot", choices=["random_dot", "textured"])
    ap.add_argument("--config", default="robotics", choices=["robotics", "c5"])
    ap.add_argument("--total-frames", type=int, default=0, help="strong scaling: total pairs split over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the second scene / C5 / latency sections")
    ap.add_argument("--cpu-frames-per-core", type=int, default=2)
    return ap.parse_args()


def config_kw(cfg):
    # C5: all five post-processing stages on both images (left/right check, small segments, gap
    # interpolation, adaptive mean, median)
    return {"filter_median": 1, "postprocess_only_left": 0} if cfg == "c5" else {}


def workload_name(a, scene=None, cfg=None):
    cfg = cfg or a.config
    return "%dx%d disp_max=%d %s ELAS + C920xK3 reproject/XR,XT/90-bin scan, %s scene" % (
        a.width, a.height, a.disp_max,
        "ROBOTICS(postprocess_only_left)" if cfg == "robotics" else "ROBOTICS+median, both images post-processed (C5)",
        scene or a.scene)


def bench_config(a, world):
    """The `config` object of the JSON line; both arms print the same one (the reference arm describes the
    bounded sample it times in cpu_baseline.sample)."""
    n = a.width * a.height
    B = a.batch
    if a.total_frames:
        pairs = "%d pairs, seeds %d..%d, split contiguously over the ranks (BASELINE config C4)" % (
            a.total_frames, SEED0, SEED0 + a.total_frames - 1)
    else:
        pairs = "%d distinct pairs per GPU, seeds %d + rank * %d + i" % (B, SEED0, B)
    return {"workload": workload_name(a), "frames_per_step_per_gpu": B, "pairs": pairs,
            "l2": "inputs per step (%.0f MB) and working set (>4 GB) exceed the 126 MB L2" % (2 * B * n / 1e6),
            "parallelism": "frames sharded across %d GPU(s), no data-path collective" % world,
            "total_frames": a.total_frames or None}


# compulsory traffic of a stage in units of N = W*H bytes per frame (ROBOTICS, left image
# post-processed: L/R 16 + segments 8 + gaps 16 + mean 16, 179 N with the 4 N of the scan; C5 post-processes both
# images and adds the median) and, for the support matcher, the integer-pipe
# ceiling (VABSDIFF4.U8.ACC issues at 16 lanes/clk/SMSP, profiles/r01_int_pipe_microbench.txt).
STAGE_BYTES_N = {"descriptor": 34.0, "support": 12.9, "dense_match": 72.0, "post": 56.0}
STAGE_BYTES_N_C5 = dict(STAGE_BYTES_N, post=128.0)     # L/R 16 + segments 16 + gaps 32 + mean 32 + median 32 (251 N in all)


def support_warp_sads(W, H, dm, step=5):
    """Warp-level VABSDIFF4 instructions the support matcher needs per frame if every lattice candidate is matched
    in both directions (elas.cpp:269-373: 4 blocks of 16 B per candidate and disparity = 16 four-byte SADs per
    lane, 32 disparities per warp instruction; ranges shorter than 10 are rejected).  An upper bound of the
    compulsory work: the backward match only runs where the forward match succeeded."""
    wc, hc = -(-W // step), -(-H // step)
    rows = sum(1 for vc in range(1, hc) if 5 <= vc * step <= H - 6)
    per_row = 0.0
    for uc in range(1, wc):
        u = uc * step
        if u < 5 or u > W - 6:
            continue
        for hi in (min(dm, u - 5), min(dm, W - u - 5)):
            if hi >= 10:
                per_row += (hi + 1) / 32.0 * 16.0
    return per_row * rows


def stage_roofline(stages_ms, W, H, dm, B, peak_gbs, sm_mhz, sms=148, cfg="robotics"):
    """max(t_HBM, t_INT) / t_measured per stage from the live stage events (ms per B-frame step), SURVEY 8(d)."""
    n = W * H
    table = STAGE_BYTES_N_C5 if cfg == "c5" else STAGE_BYTES_N
    out = {}
    for k, ms in stages_ms.items():
        if k == "raster":
            out[k] = {"bound": "scattered stores (the plane map is this implementation's intermediate: no compulsory "
                               "traffic in SURVEY 8(d))", "ms": ms, "frac": None}
            continue
        if k not in table or not ms or ms <= 0:
            out[k] = {"bound": "latency (one or two CTAs per frame)", "ms": ms, "frac": None}
            continue
        byts = table[k] * n * B
        t_hbm = byts / (peak_gbs * 1e9) * 1e3
        ent = {"bound": "hbm", "ms": ms, "algorithmic_bytes": byts, "t_hbm_ms": t_hbm, "frac": t_hbm / ms}
        if k == "support":
            # VABSDIFF4.U8.ACC: 16 lanes per clock per SM sub-partition = 0.5 warp instructions per clock, 4 per SM
            wsads = support_warp_sads(W, H, dm) * B
            t_int = wsads / (0.5 * 4 * sms * (sm_mhz or 1965.0) * 1e6) * 1e3
            ent.update({"bound": "integer pipe" if t_int > t_hbm else "hbm", "warp_sads": wsads, "t_int_ms": t_int,
                        "frac": max(t_hbm, t_int) / ms,
                        "note": "stage = matcher + inconsistency counts + support filter; warp_sads is an upper bound "
                                "of the compulsory work (every candidate matched in both directions)"})
        out[k] = ent
    return out


def q_matrix(W, H):
    import numpy as np
    import scan_lib
    fx = scan_lib.fixtures()
    return np.array(fx["Q"]["1920x1200_Kx3"] if (W, H) == (1920, 1200) else fx["Q"]["640x480"])


def digest(a):
    return hashlib.sha1(memoryview(a).cast("B")).hexdigest()


# ----------------------------------------------------------------------------- inputs
def _gen_worker(args):
    scene, W, H, dm, seed = args
    synth = importlib.import_module("jackal-navigation_b200.synth")
    I1, I2, _ = synth.SCENES[scene](W, H, dm, seed)
    return I1, I2


def make_frames(scene, W, H, dm, seeds, procs):
    """Distinct synthetic pairs, generated in parallel (1.3-1.7 s per 1920x1200 pair on one core)."""
    import numpy as np
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    jobs = [(scene, W, H, dm, s) for s in seeds]
    if procs > 1 and len(seeds) > 2:
        with mp.get_context("fork").Pool(min(procs, len(seeds))) as pool:
            for i, (a, b) in enumerate(pool.imap(_gen_worker, jobs, chunksize=1)):
                L[i], R[i] = a, b
    else:
        for i, j in enumerate(jobs):
            L[i], R[i] = _gen_worker(j)
    return L, R


# ----------------------------------------------------------------------------- CPU arm
_GATE = {}


def _cpu_worker(args):
    """One process = one core: runs the reference ELAS + the scan restatement on its frames (Triangle
    keeps mutable file-scope state, so cores are separate processes)."""
    W, H, dm, seeds, kind, scene, cfg = args
    import numpy as np
    import oracle_lib as ol
    synth = importlib.import_module("jackal-navigation_b200.synth")
    import scan_lib
    o = ol.load(kind)
    if kind == "ref":
        o.lib.ref_set_deterministic_heap(0)   # timing run: default allocator
    sp = scan_lib.ScanPort()
    fx = scan_lib.fixtures()
    Q = q_matrix(W, H)
    XR = np.array(fx["calib"]["XR"]); XT = np.array(fx["calib"]["XT"])
    pairs = {s: synth.SCENES[scene](W, H, dm, s)[:2] for s in set(seeds)}
    frames = [pairs[s] for s in seeds]
    gate = _GATE.get((W, H))       # computed once by the parent (cacheDisparityValues is init-time work)
    if gate is None:
        gate = sp.gate(Q, XR, XT, W, H)
    p = ol.robotics(dm, **config_kw(cfg))
    out = []
    per_frame = []
    t0 = time.perf_counter()
    for I1, I2 in frames:
        tf = time.perf_counter()
        D1, _ = o.process(p, I1, I2)
        r, m = sp.scan(Q, XR, XT, gate, sp.convert_u8(D1))
        per_frame.append(time.perf_counter() - tf)
        out.append((digest(D1), r.copy(), int(m.n_points)))
    return time.perf_counter() - t0, len(frames), out, per_frame


_SINGLE = {}


def cpu_arm(a, frames_per_core, cores=None, scene=None, cfg=None):
    import numpy as np
    import oracle_lib as ol
    import scan_lib
    kind = "ref" if os.path.exists(ol.REF_SO) else "port"
    cores = cores or (os.cpu_count() or 1)
    scene = scene or a.scene
    cfg = cfg or a.config
    W, H, dm = a.width, a.height, a.disp_max
    if (W, H) not in _GATE:
        fx = scan_lib.fixtures()
        _GATE[(W, H)] = scan_lib.ScanPort().gate(q_matrix(W, H), np.array(fx["calib"]["XR"]), np.array(fx["calib"]["XT"]), W, H)
    # single frame on one core, the other cores idle: median of 5 after one warm-up; once per run
    skey = (W, H, dm, kind, scene, cfg)
    if skey not in _SINGLE:
        with mp.get_context("fork").Pool(1) as pool:
            ts = pool.apply(_cpu_worker, ((W, H, dm, [5000] * 6, kind, scene, cfg),))[3]
        _SINGLE[skey] = sorted(ts[1:])[len(ts[1:]) // 2]
    t1, n1 = _SINGLE[skey], 1
    # frame k of the pool = seed SEED0 + k: the frames the GPU arm's rank 0 processes
    jobs = [(W, H, dm, [SEED0 + c * frames_per_core + k for k in range(frames_per_core)], kind, scene, cfg)
            for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    nfr = sum(r[1] for r in res)
    frames = [x for r in res for x in r[2]]      # in seed order
    return {"value": nfr / busy, "unit": UNIT, "cores": cores,
            "kind": "reference" if kind == "ref" else "port",
            "sample": "%d frames (%d per core, one process per core), max worker time %.2fs, pool wall %.2fs; "
                      "single frame on one core, others idle: %.3fs (%.3f frames/s; median of 5 after a warm-up); "
                      "ELAS = the reference's sources built -O3 -msse3; scan step (~8%% of a frame) = the plain-C "
                      "restatement with the real gate cache: the stock node indexes its 90 bins unchecked and cannot run "
                      "on this calibration (DESIGN.md section 4)" % (nfr, frames_per_core, busy, wall, t1 / n1, n1 / t1),
            "single_core_frames_per_s": n1 / t1}, frames


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cb = None
    for i in range(a.warmup + a.steps):
        cb, _ = cpu_arm(a, 1)
        if i >= a.warmup:
            vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * cb["cores"] / v, "higher_is_better": True,
            "scaling": "strong" if a.total_frames else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(a, int(os.environ.get("WORLD_SIZE", "1"))),
            "step_what": "bounded sample of the workload above: one frame per host core (%d cores) per step, "
                         "frames = seeds %d + core" % (cb["cores"], SEED0),
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


def pin_rank_to_cores(local, world):
    """One rank per GPU: give every rank its own cores on the NUMA node of its GPU (pinned buffers are then
    allocated node-locally by first touch and the copy threads of the ranks do not migrate over each other)."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus                      # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        avail = sorted(os.sched_getaffinity(0))
        cpus = avail
        if node >= 0:
            lst = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
            on_node = set()
            for part in lst.split(","):
                a, _, b = part.partition("-")
                on_node.update(range(int(a), int(b or a) + 1))
            cpus = [c for c in avail if c in on_node] or avail
        per = max(1, len(cpus) // max(1, world))
        mine = cpus[(local * per) % len(cpus):][:per] or cpus
        os.sched_setaffinity(0, mine)
        return {"numa_node": node, "cpus": "%d-%d (%d)" % (mine[0], mine[-1], len(mine))}
    except Exception as ex:      # best effort: containers without sysfs / nvidia-smi
        return {"numa_node": None, "cpus": None, "note": repr(ex)[:80]}


STAGES = ["descriptor", "support", "delaunay", "planes_grid", "raster", "dense_match", "post"]


class Runner:
    """One (scene, config) workload on this rank's GPU: device-resident and end-to-end timing."""

    def __init__(self, a, jn, torch, dist, rank, world, local, scene, cfg, B, L, R):
        import numpy as np
        self.a, self.jn, self.torch, self.dist = a, jn, torch, dist
        self.rank, self.world, self.local = rank, world, local
        self.B, self.W, self.H = B, a.width, a.height
        self.dims = (a.width, a.height, a.width)
        self.dev = torch.device("cuda", local)
        n = self.W * self.H
        self.L, self.R = L, R
        self.hL = torch.from_numpy(L).pin_memory()
        self.hR = torch.from_numpy(R).pin_memory()
        nb = L.shape[0] // B                                   # batches held by this rank
        self.nb = nb
        self.dL = self.hL.to(self.dev); self.dR = self.hR.to(self.dev)
        self.dD1 = torch.empty((B, self.H, self.W), dtype=torch.float32, device=self.dev)
        self.dStatus = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.dRanges = torch.empty((B, 90), dtype=torch.float64, device=self.dev)
        self.dMeta = torch.empty((B, 5), dtype=torch.float64, device=self.dev)   # jn_scan_meta = 40 bytes
        self.dU8 = torch.empty((B, self.H, self.W), dtype=torch.uint8, device=self.dev)
        # second output set: two rolling submissions are in flight at a time
        self.dOut = [(self.dD1, self.dStatus, self.dRanges, self.dMeta, self.dU8),
                     (torch.empty_like(self.dD1), torch.zeros_like(self.dStatus), torch.empty_like(self.dRanges),
                      torch.empty_like(self.dMeta), torch.empty_like(self.dU8))]
        # host result buffers of the end-to-end path, one set per in-flight submission
        self.hRanges = [torch.empty((B, 90), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.hMeta = [torch.empty((B, 5), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.hStatus = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.hU8 = [torch.empty((B, self.H, self.W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.elas = jn.Elas(jn.parameters(jn.ROBOTICS, disp_max=a.disp_max, **config_kw(cfg)), device=local)
        import scan_lib
        cal = jn.Calibration(scan_lib.CALIB_YML)
        cal.set_q_matrix(q_matrix(self.W, self.H))
        self.scan = jn.ObstacleScan(cal, self.W, self.H, device=local)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.h2d = 2 * B * n
        self.d2h = B * (n + 90 * 8 + 40 + 4)
        self.it = 0

    def step_resident(self, k=0):
        B, n = self.B, self.W * self.H
        self.elas.process_batch(self.dL.data_ptr() + k * B * n, self.dR.data_ptr() + k * B * n, self.dD1.data_ptr(), 0,
                                self.dStatus.data_ptr(), self.dims, B, self.stream.cuda_stream)
        self.scan.from_disparity_batch(B, self.dD1.data_ptr(), self.dRanges.data_ptr(), self.dMeta.data_ptr(),
                                       self.dU8.data_ptr(), self.stream.cuda_stream)

    def step_resident_rolling(self, k=0):
        """Device-resident frames through jn_stereo_scan_submit_device: the same kernels as step_resident, queued
        on the library's rolling sub-batch streams (no fork/join on a caller stream between steps)."""
        B, n = self.B, self.W * self.H
        D1, st, rg, mt, u8 = self.dOut[self.it & 1]
        self.it += 1
        self.elas.stereo_scan_submit_device(self.scan, B, self.dL.data_ptr() + k * B * n, self.dR.data_ptr() + k * B * n,
                                            self.dims, D1.data_ptr(), st.data_ptr(), rg.data_ptr(), mt.data_ptr(),
                                            u8.data_ptr())

    def step_e2e_scans(self, k=0):
        """Same call, the optional u8 disparity maps not requested: scans, meta and status come back."""
        B, n = self.B, self.W * self.H
        j = self.it & 1
        self.it += 1
        self.elas.stereo_scan_submit(self.scan, B, self.hL.data_ptr() + k * B * n, self.hR.data_ptr() + k * B * n,
                                     self.dims, self.hRanges[j].data_ptr(), self.hMeta[j].data_ptr(),
                                     self.hStatus[j].data_ptr())

    def step_e2e(self, k=0):
        """The call a user of the C ABI makes: host image pairs in, scans + u8 maps out (asynchronous;
        up to two submissions overlap inside the library)."""
        B, n = self.B, self.W * self.H
        j = self.it & 1
        self.it += 1
        self.elas.stereo_scan_submit(self.scan, B, self.hL.data_ptr() + k * B * n, self.hR.data_ptr() + k * B * n,
                                     self.dims, self.hRanges[j].data_ptr(), self.hMeta[j].data_ptr(),
                                     self.hStatus[j].data_ptr(), self.hU8[j].data_ptr())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        torch = self.torch
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for s in range(steps):
                fn(s % self.nb)
            if finish:
                finish()                    # host-blocking: every queued batch has landed
            e1.record(self.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def measure(self, steps, warmup):
        jn, torch = self.jn, self.torch
        with torch.cuda.stream(self.stream):
            for s in range(max(warmup, 3)):
                self.step_resident(s % self.nb)
        torch.cuda.synchronize()
        self.ms_forkjoin = self.timed(self.step_resident, steps)     # stream API: fork/join on the caller's stream
        for s in range(max(warmup, 3)):
            self.step_resident_rolling(s % self.nb)
        self.elas.stereo_scan_wait()
        l0 = jn.launch_count()
        ms = self.timed(self.step_resident_rolling, steps, self.elas.stereo_scan_wait)
        launches = jn.launch_count() - l0
        for s in range(3):
            self.step_e2e(s % self.nb)
        self.elas.stereo_scan_wait()
        ms_e2e = self.timed(self.step_e2e, steps, self.elas.stereo_scan_wait)
        self.ms_e2e_scans = self.timed(self.step_e2e_scans, steps, self.elas.stereo_scan_wait)
        return ms, ms_e2e, launches

    def stage_times(self, reps=3):
        import numpy as np
        lib = self.jn.lib()
        lib.jn_elas_profile.argtypes = [C.c_void_p, C.c_int]
        lib.jn_elas_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        lib.jn_elas_profile(self.elas._h, 1)
        acc = np.zeros(7, np.float64)
        for _ in range(reps):
            with self.torch.cuda.stream(self.stream):
                self.step_resident(0)
            self.torch.cuda.synchronize()
            buf = (C.c_float * 7)()
            lib.jn_elas_profile_read(self.elas._h, buf)
            acc += np.array(list(buf))
        lib.jn_elas_profile(self.elas._h, 0)
        return acc / reps

    def close(self):
        self.elas.close(); self.scan.close()


def latency_section(a, jn, torch, local, L, R):
    """The reference-shaped single-frame call (Elas::process, host buffers, synchronous)."""
    import numpy as np
    W, H, dm = a.width, a.height, a.disp_max
    dims = (W, H, W)
    out = {}

    def med(fn, reps=12):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        ts = sorted(ts[2:])
        return 1000.0 * ts[len(ts) // 2]

    p = jn.parameters(jn.ROBOTICS, disp_max=dm)
    e1 = jn.Elas(p, device=local)
    D1 = np.zeros((H, W), np.float32); D2 = np.zeros((H, W), np.float32)
    out["pageable_D1_D2_ms"] = med(lambda: e1.process(L[0], R[0], D1, D2, dims))
    out["pageable_D1_only_ms"] = med(lambda: e1.process(L[0], R[0], D1, None, dims))
    pL = torch.from_numpy(L[0].copy()).pin_memory().numpy(); pR = torch.from_numpy(R[0].copy()).pin_memory().numpy()
    pD1 = torch.zeros((H, W)).pin_memory().numpy()
    out["pinned_D1_only_ms"] = med(lambda: e1.process(pL, pR, pD1, None, dims))
    e1.close()

    def fresh():
        e = jn.Elas(p, device=local)                # point_cloud.cpp:416-419: a new Elas for every frame
        e.process(pL, pR, pD1, None, dims)
        e.close()
    out["fresh_elas_per_frame_pinned_D1_only_ms"] = med(fresh)
    # per-stage device times of a one-frame batch
    lib = jn.lib()
    e2 = jn.Elas(p, device=local)
    dL = torch.from_numpy(L[:1]).cuda(local); dR = torch.from_numpy(R[:1]).cuda(local)
    dD = torch.empty((1, H, W), dtype=torch.float32, device=dL.device); dS = torch.zeros(1, dtype=torch.int32, device=dL.device)
    lib.jn_elas_profile(e2._h, 1)
    acc = np.zeros(7)
    for i in range(5):
        e2.process_batch(dL.data_ptr(), dR.data_ptr(), dD.data_ptr(), 0, dS.data_ptr(), dims, 1, 0)
        torch.cuda.synchronize()
        buf = (C.c_float * 7)(); lib.jn_elas_profile_read(e2._h, buf)
        if i >= 2:
            acc += np.array(list(buf))
    lib.jn_elas_profile(e2._h, 0)
    e2.close()
    out["stage_ms_one_frame"] = {k: float(v / 3) for k, v in zip(STAGES, acc)}
    return out


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    jn = importlib.import_module("jackal-navigation_b200")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference)")
    W, H, dm, B = a.width, a.height, a.disp_max, a.batch
    n = W * H
    procs = max(1, (os.cpu_count() or 1) // world)

    # ---- synthetic inputs (worker processes are forked before CUDA / NCCL are initialised)
    if a.total_frames:                       # C4: contiguous shard of seeds SEED0 .. SEED0 + T - 1
        sharding = importlib.import_module("jackal-navigation_b200.sharding")
        lo, hi = sharding.frame_shard(a.total_frames, rank, world)
        cnt = ((hi - lo) // B) * B
        if cnt == 0:
            raise SystemExit("--total-frames: every rank needs at least one batch of %d frames" % B)
        seeds = [SEED0 + lo + i for i in range(cnt)]
    else:                                    # B distinct pairs per rank
        seeds = [SEED0 + rank * B + i for i in range(B)]
    t_gen = time.perf_counter()
    L, R = make_frames(a.scene, W, H, dm, seeds, procs)
    t_gen = time.perf_counter() - t_gen
    # second scene / C5 / single-frame latency: sections of the single-GPU line only (the scaling runs measure the
    # headline workload and nothing else, exactly the code path of the N > 1 records under profiles/)
    extras_on = not a.no_extras and not a.total_frames and world == 1
    other_scene = "textured" if a.scene == "random_dot" else "random_dot"
    nd2 = min(16, B)
    if extras_on:
        l2, r2 = make_frames(other_scene, W, H, dm, [SEED0 + rank * B + i for i in range(nd2)], procs)
        L2o = np.concatenate([l2] * (B // nd2)); R2o = np.concatenate([r2] * (B // nd2))

    pin = pin_rank_to_cores(local, world) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    run = Runner(a, jn, torch, dist, rank, world, local, a.scene, a.config, B, L, R)
    sampler = ClockSampler(local) if rank == 0 else None
    steps = a.steps if not a.total_frames else run.nb      # strong scaling: one pass over the shard
    ms, ms_e2e, launches = run.measure(steps, a.warmup)
    clocks = sampler.stop() if sampler else None            # sampled over both timed regions
    frames_done = B * steps if not a.total_frames else a.total_frames // world // B * B
    value = world * frames_done / (ms / 1000.0)
    e2e = world * frames_done / (ms_e2e / 1000.0)
    e2e_scans = world * frames_done / (run.ms_e2e_scans / 1000.0)
    stages = run.stage_times()

    # ---- outputs of the first batch for the parity check (device-resident path and host path)
    with torch.cuda.stream(run.stream):
        run.step_resident(0)
    torch.cuda.synchronize()
    status = run.dStatus.cpu().numpy()
    nchk = min(B, (os.cpu_count() or 1) * a.cpu_frames_per_core, 32)
    gD1 = run.dD1[:nchk].cpu().numpy()
    gRanges = run.dRanges[:nchk].cpu().numpy()
    run.it = 0
    run.step_e2e(0); run.elas.stereo_scan_wait()
    hostRanges = run.hRanges[0][:nchk].numpy().copy()

    extras = {}
    if rank == 0 and extras_on:
        extras["single_frame"] = latency_section(a, jn, torch, local, L, R)
    run.close()

    # ---- second scene family and BASELINE config C5: short runs of the same measurement
    if extras_on:
        others = [("scenes", other_scene, a.config)]
        if a.config != "c5":
            others.append(("c5", a.scene, "c5"))
        for key, scene, cfg in others:
            nd = nd2 if scene != a.scene else B      # distinct pairs (the rest of the batch repeats them)
            La, Ra = (L, R) if scene == a.scene else (L2o, R2o)
            r2 = Runner(a, jn, torch, dist, rank, world, local, scene, cfg, B, La, Ra)
            m2, m2e, _ = r2.measure(10, 3)
            st2 = r2.stage_times(2)
            ent = {"workload": workload_name(a, scene, cfg), "value": world * B * 10 / (m2 / 1000.0), "unit": UNIT,
                   "e2e": world * B * 10 / (r2.ms_e2e_scans / 1000.0), "e2e_with_u8_maps": world * B * 10 / (m2e / 1000.0),
                   "ms_per_step": m2 / 10, "steps": 10,
                   "distinct_pairs_per_gpu": nd,
                   "stage_ms_per_step": {k: float(v) for k, v in zip(STAGES, st2)},
                   "frames_ok": int((r2.dStatus.cpu().numpy() == 0).sum())}
            if cfg == "c5":
                ent["algorithmic_bytes_per_frame"] = 251 * n
            extras[key] = ent if key == "c5" else {scene: ent}
            r2.close()
            del r2

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
    tsrc = None
    tpath = os.path.join(ROOT, "profiles", "dense_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_frame") * B   # per launch of B frames, like `achieved`
            tsrc = tj.get("source")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if a.total_frames else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(a, world),
        "input_generation_s": t_gen,
        "mpix_per_s": value * n / 1e6, "ms_per_frame": ms / steps / B,
        "value_what": "device-resident image pairs through jn_stereo_scan_submit_device / _wait (ELAS + scan kernels on "
                      "the library's rolling sub-batch streams, two submissions in flight, outputs stay on the device)",
        "stream_api": {"value": world * frames_done / (run.ms_forkjoin / 1000.0), "unit": UNIT,
                       "ms_per_step": run.ms_forkjoin / steps,
                       "what": "same frames through jn_elas_process_batch + jn_scan_from_disparity_batch on one caller "
                               "stream (sub-batches fork from and join back into it every step)"},
        "frames_ok": int((status == 0).sum()), "frames_few_support": int((status == 1).sum()),
        "e2e": {"value": e2e_scans, "unit": UNIT, "h2d_bytes_per_step": run.h2d, "d2h_bytes_per_step": B * (90 * 8 + 40 + 4),
                "ms_per_step": run.ms_e2e_scans / steps,
                "what": "C ABI jn_stereo_scan_submit/_wait: pinned host image pairs -> H2D -> ELAS + scan kernels -> "
                        "D2H of the step's result (90-bin scans, scan meta, status per frame) into pinned host buffers; "
                        "copies on the library's own streams, double-buffered, all inside the timed region",
                "with_u8_maps": {"value": e2e, "unit": UNIT, "d2h_bytes_per_step": run.d2h, "ms_per_step": ms_e2e / steps,
                                 "what": "same call with the optional convertTo(CV_8U) disparity maps (the reference "
                                         "publishes them for display) also copied back: +W*H bytes per frame; on an "
                                         "8-GPU box this variant sits on the host's DMA ceiling (profiles/r02_copy_ceiling.txt)"},
                "rank_pinning": pin},
        "gpu_launches": int(launches),
        "stage_ms_per_step": {k: float(v) for k, v in zip(STAGES, stages)},
        "stage_roofline": stage_roofline({k: float(v) for k, v in zip(STAGES, stages)}, W, H, dm, B, peak,
                                         (clocks or {}).get("sm_mhz"), cfg=a.config),
        "roofline": {"bound": "hbm", "kernel": "dense_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "algorithmic_bytes_per_launch": 72.0 * n * B, "kernel_ms": dense_ms},
        "clocks": clocks,
    }
    line.update(extras)
    if "single_frame" in extras:
        line["single_frame_latency_ms"] = extras["single_frame"]["pinned_D1_only_ms"]
    if not a.no_cpu_baseline and world == 1:
        try:
            cb, frames = cpu_arm(a, a.cpu_frames_per_core)
            line["cpu_baseline"] = cb
            # ---- parity inside the bench: the CPU checker re-computes frames of the timed batch
            k = min(nchk, len(frames))
            d1_ok = sum(digest(np.ascontiguousarray(gD1[i])) == frames[i][0] for i in range(k))
            occ_ok = sum(bool(np.array_equal(gRanges[i] < 1e9 - 1, frames[i][1] < 1e9 - 1)) for i in range(k))
            rng_err = max(float(np.abs(np.where(gRanges[i] < 1e9 - 1, gRanges[i] - frames[i][1], 0)).max()) for i in range(k))
            host_ok = sum(bool(np.array_equal(hostRanges[i], gRanges[i])) for i in range(k))
            line["parity"] = {"frames_checked": k, "checker": cb["kind"], "d1_maps_bit_equal": d1_ok,
                              "scan_bins_equal": occ_ok, "scan_range_max_abs_err_m": rng_err,
                              "host_path_equals_device_path": host_ok,
                              "ok": bool(d1_ok == k and occ_ok == k and rng_err <= 1e-9 and host_ok == k)}
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
