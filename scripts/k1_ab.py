#!/usr/bin/env python
"""A/B of kernel builds: for every library given, one process (LQ_LIB) runs square 1024 x 1024 at beta
(default 128: the headline tiling and page size at 14 GB), `therm` untimed steps from the empty state, then
`steps` steps with the phase timers on; prints ms per step of every phase and a checksum of the trajectory
(operators and clusters per step: builds that only differ in code generation must agree).
usage: k1_ab.py [--beta B] [--therm T] [--steps S] lib.so [lib.so ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np
import looper_b200 as lq
beta, therm, steps = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
eng = lq.Engine(lq.hypercubic_lattice((1024, 1024)), beta, seed=29833, tile_sites=256, window_ops=3.0)
eng.sweep_many(therm, collect=False)
eng.enable_timers(True)
out = eng.sweep_many(steps)
t = {x["label"] or str(x["id"]): round(1e3 * x["seconds"] / steps, 4) for x in eng.timers() if x["count"]}
print(json.dumps(dict(nop=float(out["nop"].sum()), nc=float(out["nc"].sum()), ms=t, total=round(sum(t.values()), 3))))
eng.close()
""" % ROOT

if __name__ == "__main__":
    args = sys.argv[1:]
    beta, therm, steps = 128.0, 12, 8
    while args and args[0].startswith("--"):
        k, v = args[0], args[1]
        args = args[2:]
        if k == "--beta": beta = float(v)
        elif k == "--therm": therm = int(v)
        elif k == "--steps": steps = int(v)
    for lib in args:
        env = dict(os.environ, LQ_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CHILD, str(beta), str(therm), str(steps)], env=env,
                           capture_output=True, text=True, timeout=120)
        print(os.path.basename(lib), r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
