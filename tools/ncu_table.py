"""One line per kernel from a multi-kernel `ncu --set full` report (first captured launch of every kernel name).
    python tools/ncu_table.py rep.ncu-rep [frames_per_launch] > profiles/rNN_ncu_kernels.txt
Columns: duration, DRAM bytes (read + write), DRAM throughput % of peak, issue-slot utilisation, ALU / FMA / LSU
pipe utilisation, L1 data-pipe (lsu wavefronts) %, warps active %, registers, the two largest stall reasons."""
import csv, io, re, subprocess, sys
path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def col(r, name, scale=1.0):
    i = ix.get(name)
    if i is None or r[i] in ("", "n/a"): return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i]
    if name.startswith("dram__bytes") or name.startswith("lts__t_bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    return v * scale
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
seen = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
    if name in seen: continue
    seen[name] = r
print("%-34s %9s %9s %6s %6s %5s %5s %5s %5s %6s %4s  %s" % ("kernel", "us", "DRAM MB", "DRAM%", "issue%", "ALU%", "FMA%", "LSU%", "L1d%", "warps%", "regs", "top stalls (cycles per issue)"))
for name, r in seen.items():
    st = sorted(((col(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls if h.find("selected") < 0), reverse=True)[:2]
    print("%-34s %9.1f %9.1f %6.1f %6.1f %5.1f %5.1f %5.1f %5.1f %6.1f %4d  %s" % (
        name[:34], col(r, "gpu__time_duration.sum"), (col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")) / 1e6,
        col(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + col(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"), col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        col(r, "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"), col(r, "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
        col(r, "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"), col(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(col(r, "launch__registers_per_thread")),
        ", ".join("%s %.1f" % (n, v) for v, n in st)))
