#!/usr/bin/env python
"""profiles/traffic.json and a per-kernel table from an ncu launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (csv log file).
usage: make_traffic.py <launches.csv> <workload> <operators per step> "<ncu command>" [peak GB/s]"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path, workload, nop, how = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
peak = float(sys.argv[5]) if len(sys.argv) > 5 else 6548.5
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
d = defaultdict(lambda: defaultdict(list))
for r in rows[1:]:
    name = re.sub(r"<.*", "", r[ki].split("(")[0].replace("void ", "")).strip()
    d[name][r[mi]].append(float(r[vi].replace(",", "")))
mean = lambda x: sum(x) / len(x)
kern = {}
for k, v in d.items():
    kern[k] = {"dram_bytes_per_launch": mean(v["dram__bytes_read.sum"]) + mean(v["dram__bytes_write.sum"]),
               "ncu_duration_ms": mean(v["gpu__time_duration.sum"]) / 1e6, "launches": len(v["gpu__time_duration.sum"])}
kern = dict(sorted(kern.items(), key=lambda kv: -kv[1]["ncu_duration_ms"]))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump({"workload": workload, "n_gpus": 1, "operators_per_step": nop, "how": how + " (" + os.path.basename(path) + "); bytes per launch",
           "kernels": kern}, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
tot_ms = sum(k["ncu_duration_ms"] for k in kern.values())
tot_b = sum(k["dram_bytes_per_launch"] for k in kern.values())
print("| kernel | ms per launch | share | DRAM GB per launch | B/op | GB/s |")
print("|---|---|---|---|---|---|")
for k, v in kern.items():
    gb = v["dram_bytes_per_launch"] / 1e9
    print(f"| {k} | {v['ncu_duration_ms']:.2f} | {100 * v['ncu_duration_ms'] / tot_ms:.1f}% | {gb:.1f} | "
          f"{v['dram_bytes_per_launch'] / nop:.1f} | {gb / v['ncu_duration_ms'] * 1e3:.0f} |")
print(f"| sum | {tot_ms:.1f} | | {tot_b / 1e9:.0f} | {tot_b / nop:.0f} | {tot_b / 1e6 / tot_ms:.0f} |")
print(f"\nAt the measured copy bandwidth ({peak:.0f} GB/s) this traffic would take {tot_b / 1e6 / peak:.1f} ms.")
