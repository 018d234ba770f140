#!/usr/bin/env python
"""Per-kernel durations of the LAST `steps` Monte Carlo steps of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`): the list is cut at the launches of
k_diag_update.  usage: launch_means.py launches.csv [steps] [more.csv ...] -- several files are shown side by side."""
import csv
import sys
from collections import OrderedDict


def load(path, steps):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    seq = []
    for r in rows[rows.index(hdr) + 1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        seq.append((r[ki].split("(")[0].replace("void ", ""), v * scale))
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_diag_update")]
    if len(starts) <= steps:
        raise SystemExit(f"{path}: only {len(starts)} steps captured")
    lo, hi = starts[-steps - 1], starts[-1]   # the last complete `steps` steps
    out = OrderedDict()
    for k, v in seq[lo:hi]:
        out[k] = out.get(k, 0.0) + v / steps
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    steps = 2
    files = []
    for a in args:
        if a.isdigit():
            steps = int(a)
        else:
            files.append(a)
    tabs = [load(f, steps) for f in files]
    names = list(OrderedDict.fromkeys(k for t in tabs for k in t))
    print("| kernel | " + " | ".join(f.split("/")[-1] for f in files) + " |")
    print("|---|" + "---|" * len(files))
    for k in names:
        print(f"| {k} | " + " | ".join(f"{t.get(k, 0.0):.1f}" for t in tabs) + " |")
    print("| total (us per step) | " + " | ".join(f"{sum(t.values()):.1f}" for t in tabs) + " |")
