#!/usr/bin/env python
"""Small driver for ncu captures: square L x L at beta, `therm` steps untimed, then `steps` steps.
usage: ncu_probe.py L beta therm steps [tile_sites [window_ops]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import looper_b200 as lq  # noqa: E402

L, beta, therm, steps = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
tile = int(sys.argv[5]) if len(sys.argv) > 5 else 0
wops = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
eng = lq.Engine(lq.hypercubic_lattice((L, L)), beta, seed=29833, tile_sites=tile, window_ops=wops)
eng.sweep_many(therm, collect=False)
out = eng.sweep_many(steps)
print("nop", float(out["nop"].mean()), "nc", float(out["nc"].mean()))
eng.close()
