"""SASS opcode histogram of every kernel of this repository (cuobjdump -sass on the object files), plus the markers the profiling
recipe asks for: VABSDIFF4 (byte SAD on the integer pipe), UBLKCP / SYNCS (1-D bulk async copies + mbarrier =
TMA unit), REDUX, MATCH, no HMMA / UTCMMA (nothing on this path is a contraction).
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objdir = os.path.join(ROOT, "jackal-navigation_b200", "_obj")     # our translation units (the .so also holds nvJPEG's)
out = "".join(subprocess.run(["cuobjdump", "-sass", os.path.join(objdir, f)], capture_output=True, text=True).stdout
              for f in sorted(os.listdir(objdir)) if f.endswith(".o"))
per = collections.OrderedDict(); cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\((?!anonymous).*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        cur = per.setdefault(name, collections.Counter()); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        cur[".".join(m.group(1).split(".")[:2])] += 1
tot = collections.Counter()
for c in per.values(): tot.update(c)
mark = ["VABSDIFF4.U8", "UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "REDUX", "MATCH.ANY", "HMMA", "UTCHMMA", "UTMALDG", "ATOMG", "RED", "CCTL"]
print("kernels: %d, SASS instructions: %d" % (len(per), sum(tot.values())))
print("markers over all kernels: " + ", ".join("%s %d" % (m, sum(v for k, v in tot.items() if k.startswith(m))) for m in mark))
for name, c in per.items():
    n = sum(c.values())
    print("\n%s  (%d instructions)" % (name, n))
    print("   " + "  ".join("%s %d" % kv for kv in c.most_common(14)))
