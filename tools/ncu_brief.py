"""Prints the handful of ncu metrics we steer by from a .ncu-rep (raw page CSV)."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp", "smsp__warps_issue_stalled", "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_uniform", "l1tex__throughput",
        "smsp__thread_inst_executed_per_inst_executed", "sm__pipe_fmaheavy", "sm__pipe_fma_cycles", "launch__grid_size", "launch__block_size",
        "smsp__pcsamp_warps_issue_stalled"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", path)
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) for k in KEYS):
            if "pcsamp" in h and v in ("0", ""): continue
            print("  %-90s %s %s" % (h, v, u))
