#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,...]` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        metric, unit = r["Metric Name"], r["Metric Unit"]
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        d = per.setdefault(name, collections.defaultdict(float))
        if metric == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
            d["ms"] += v * scale
            d["n"] += 1
        elif metric.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
            d["dram_bytes"] += v * scale
        elif "pipe_tensor" in metric:
            d["tensor_pct_sum"] += v
    total = sum(d["ms"] for d in per.values())
    print(f"{'ms':>10} {'share':>6} {'n':>5} {'ms/launch':>10} {'DRAM GB/launch':>15} {'tensor%':>8}  kernel")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"])[:30]:
        n = max(d["n"], 1)
        print(f"{d['ms']:10.3f} {100 * d['ms'] / total:5.1f}% {int(n):5d} {d['ms'] / n:10.3f} "
              f"{d['dram_bytes'] / n / 1e9:15.3f} {d['tensor_pct_sum'] / n:8.1f}  {name[:90]}")
    print(f"{total:10.3f} ms total over {int(sum(d['n'] for d in per.values()))} launches")


if __name__ == "__main__":
    main(sys.argv[1])
