#!/usr/bin/env python
"""Join ncu's per-SASS-instruction counters with source lines (nvdisasm --print-line-info) and
print the hottest source lines of one kernel.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <mangled function name> [liblq.so]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kre, mangled = sys.argv[1:4]
so = sys.argv[4] if len(sys.argv) > 4 else "alps-looper_b200/liblq.so"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
ii, ti, wi = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
sass = []
for r in rows[hi + 1:]:
    if len(r) <= ii or not r[0].startswith("0x"):
        if sass: break          # second captured instance of the kernel starts here
        continue
    sass.append((r[1].strip(), float(r[ii] or 0), float(r[ti] or 0), float(r[wi] or 0)))
# stop at second kernel instance if several were captured
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":"))
cur = None
ins = []
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"):
        if ins: break
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.+?);", l)
    if m:
        ins.append((cur, m.group(1).strip()))
n = min(len(ins), len(sass))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for k in range(n):
    a = agg[ins[k][0]]
    a[0] += sass[k][1]; a[1] += sass[k][2]; a[2] += sass[k][3]
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
print(f"kernel {kre}: {len(sass)} SASS rows, {len(ins)} disassembled; total warp-inst {tot:.3g}")
src_cache = {}
def src(fl):
    if fl is None: return "?"
    f, ln = fl
    for base in ("alps-looper_b200/csrc/", ""):
        pth = base + f
        if os.path.exists(pth):
            if pth not in src_cache: src_cache[pth] = open(pth).read().splitlines()
            L = src_cache[pth]
            return L[ln - 1].strip()[:100] if ln - 1 < len(L) else ""
    return ""
for fl, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOPN","32"))]:
    print(f"{100*v[0]/tot:5.1f}% inst  {100*v[2]/tots:5.1f}% stall  lanes={v[1]/max(v[0],1):4.1f}  {fl}: {src(fl)}")
