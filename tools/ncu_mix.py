"""Instruction mix + hottest SASS lines of a kernel from `ncu --page source --csv` output of a .ncu-rep."""
import csv, subprocess, sys, io, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
S, E, SMP = ix["Source"], ix["Instructions Executed"], ix["# Samples"]
mix = collections.Counter(); tot = 0; stot = 0; lines = []
for r in rows[2:]:
    try: n = int(r[E]); s = int(r[SMP])
    except Exception: continue
    op = r[S].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    mix[".".join(o.split(".")[:3])] += n; tot += n; stot += s
    lines.append((s, n, r[S].strip()))
print("warp instructions executed:", tot, " samples:", stot)
for k, v in mix.most_common(top): print("  %-28s %12d %5.1f%%" % (k, v, 100.0 * v / tot))
print("hottest by stall samples:")
for s, n, t in sorted(lines, reverse=True)[:top]: print("  %6d %5.1f%% x%-10d %s" % (s, 100.0 * s / max(stot, 1), n, t[:100]))
