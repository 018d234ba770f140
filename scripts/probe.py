import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import looper_b200 as lq

def run(L, beta, therm, steps, tile=0, wops=0.0, timers=True):
    lat = lq.hypercubic_lattice((L, L))
    t0 = time.perf_counter()
    eng = lq.Engine(lat, beta, tile_sites=tile, window_ops=wops, timers=False)
    info = eng.info()
    t1 = time.perf_counter()
    eng.sweep_many(therm, collect=False)
    t2 = time.perf_counter()
    out = eng.sweep_many(steps)
    t3 = time.perf_counter()
    nop = out["nop"].mean()
    print(json.dumps(dict(L=L, beta=beta, tile=tile, wops=wops, info=info, create_s=t1-t0, therm_s=t2-t1,
          ms_per_mcs=1e3*(t3-t2)/steps, nop=nop, nc=out["nc"].mean(), gops=nop*steps/(t3-t2)/1e9,
          frac=nop*steps/(t3-t2)*104/6534.1e9)))
    eng.close()
    if timers:
        eng = lq.Engine(lat, beta, tile_sites=tile, window_ops=wops, timers=True)
        eng.sweep_many(therm, collect=False)
        t = eng.timers()
        tot = sum(x["seconds"] for x in t)
        print("  phases:", ", ".join("%s=%.1f%%" % (x["label"], 100*x["seconds"]/tot) for x in t), " ms/mcs(sum)=%.3f" % (1e3*tot/therm))
        eng.close()

if __name__ == "__main__":
    L = int(sys.argv[1]); beta = float(sys.argv[2]); therm = int(sys.argv[3]); steps = int(sys.argv[4])
    tile = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    wops = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    run(L, beta, therm, steps, tile, wops)
