#!/usr/bin/env python
"""Turn gpurun_out/launches_*.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/prof_*.ncu-rep (ncu --set full) into the text summaries kept under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out = ["| kernel | launches | mean us | share of captured time |", "|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled")
             and not h.endswith("_not_issued")]
    ki = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        out.append(f"### {r[ki].split('(')[0]}")
        for i in idx:
            out.append(f"- {hdr[i]} = {r[i]} {units[i]}")
        vals = sorted(((hdr[i].replace('smsp__pcsamp_warps_issue_stalled_', ''),
                        float(r[i].replace(',', '') or 0)) for i in stall), key=lambda x: -x[1])
        tot = sum(v for _, v in vals) or 1
        out.append("- stall reasons: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in vals[:5]))
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(launches(path) if kind == "launches" else full(path))
