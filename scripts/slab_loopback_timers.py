"""P slab engines of ONE Markov chain on one GPU through the loopback communicator (for `ncu
--metrics gpu__time_duration.sum`: per-kernel durations of the slab-only kernels; the stream timers
of the engines are useless here because the engines share the GPU).
usage: slab_loopback_timers.py [P] [L] [beta] [steps]"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import looper_b200 as lq   # noqa: E402

spec = importlib.util.spec_from_file_location("lq_comm", os.path.join(ROOT, "alps-looper_b200", "comm.py"))
comm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(comm)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
beta = float(sys.argv[3]) if len(sys.argv) > 3 else 128.0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
lat = lq.hypercubic_lattice((L, L))
grp = comm.LoopbackGroup(P)


def body(r):
    eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=29833, tile_sites=256, reserve=1.4)
    grp.attach(eng, r)
    out = eng.sweep_many(steps)
    eng.close()
    return float(out["nop"][-1]), float(out["noc"][-1])


print(grp.run(body)[0])
