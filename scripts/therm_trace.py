"""Stationarity trace of the bench workload: nop, nc, staggered magnetisation^2 per Monte Carlo step from
the empty state out to 2 beta steps (SURVEY 8d asks for max(64, 2 beta) thermalisation steps; bench.py
uses 64 -- this trace shows what the operator / cluster counts, which are what the kernels' work
depends on, do after that).  usage: therm_trace.py [workload] [steps] > profiles/r02_thermalisation_trace.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import looper_b200 as lq   # noqa: E402
import bench               # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "square1024_beta1024"
dims, beta, therm, sse = bench.workload(name)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else int(2 * beta)
lat = lq.hypercubic_lattice(dims)
eng = lq.Engine(lat, beta, seed=29833, tile_sites=256 if lat["num_sites"] >= 512 * 512 else 64, reserve=1.4)
out = []
done = 0
while done < steps:
    m = min(64, steps - done)
    out.append(eng.sweep_many(m))
    done += m
out = np.concatenate(out)
N = lat["num_sites"]
print(f"# Thermalisation trace, {name} (seed 29833, from the empty all-up state), {steps} steps\n")
print("| step | operators | clusters | E/N | staggered magnetisation^2 / N^2 |")
print("|---|---|---|---|---|")
marks = sorted(set([1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, steps]))
for s in marks:
    if s <= steps:
        r = out[s - 1]
        print(f"| {s} | {r['nop']:.0f} | {r['nc']:.0f} | {r['ene'] / N:.5f} | {r['smag2'] / N / N:.5f} |")
def seg(a, b):
    x = out[a:b]
    return f"operators {x['nop'].mean():.4e} (rms {x['nop'].std() / x['nop'].mean():.1e}), clusters {x['nc'].mean():.4e}, E/N {x['ene'].mean() / N:.5f}, ms^2/N^2 {x['smag2'].mean() / N / N:.5f}"
print()
print(f"* steps 65-128: {seg(64, 128)}")
if steps >= 1024:
    print(f"* steps 513-1024: {seg(512, 1024)}")
print(f"* last quarter ({steps - steps // 4 + 1}-{steps}): {seg(steps - steps // 4, steps)}")
eng.close()
