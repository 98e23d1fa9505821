"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/summarize_launches.py profiles/r01_launches_first.csv [launches_of_each_kernel_per_step]
"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        t *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        per.setdefault(name, []).append(t)
    steps = len(per.get("dense_kernel", [1]))
    tot = sum(sum(v) for v in per.values()) / steps / 1e6
    print("steps captured: %d   sum of kernel times: %.2f ms/step (cold-cache, serialised)" % (steps, tot))
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        ms = sum(v) / steps / 1e6
        print("%8.3f ms/step %5.1f%%  x%-4.1f %s" % (ms, 100 * ms / tot, len(v) / steps, k))


if __name__ == "__main__":
    main(sys.argv[1])
