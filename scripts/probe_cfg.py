import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import looper_b200 as lq

def run(name, lat, beta, weights, therm, steps, **kw):
    eng = lq.Engine(lat, beta, weights=weights, **kw)
    info = eng.info()
    eng.sweep_many(therm, collect=False)
    t0 = time.perf_counter(); out = eng.sweep_many(steps); t1 = time.perf_counter()
    nop = out["nop"].mean()
    print(json.dumps(dict(cfg=name, ms_per_mcs=1e3*(t1-t0)/steps, nop=nop, nc=out["nc"].mean(), gops=nop*steps/(t1-t0)/1e9,
          ene_density=out["ene"].mean()/lat["num_sites"], tiles=info["num_tiles"], W=info["num_windows"], cap=info["page_capacity"], npo=info["nodes_per_op"], GB=info["device_bytes"]/1e9)))
    eng.close()

v_xxz, _, _ = lq.xxz_weights(1.0, 0.5)
run("config5(i) chain L=4096 beta=256 XXZ Jz=0.5", lq.chain_lattice(4096), 256.0, tuple(v_xxz), 300, 50)
v_tfi, _, _ = lq.xxz_weights(0.0, 1.0)
run("config5(ii) chain L=4096 beta=256 TFI Jz=1 Gamma=0.7", lq.chain_lattice(4096), 256.0, tuple(v_tfi), 300, 50, site_weight=0.35)
run("config1 chain L=16 T=0.1", lq.chain_lattice(16), 10.0, (0.5,0,0,0), 300, 200)
run("config4 cubic 64^3 T=0.95 (path integral)", lq.hypercubic_lattice((64,64,64)), 1/0.95, (0.5,0,0,0), 100, 50)
run("config2 square 256 beta=64 (tile 256)", lq.hypercubic_lattice((256,256)), 64.0, (0.5,0,0,0), 200, 50, tile_sites=256)
run("config2 square 256 beta=64 (tile 64)", lq.hypercubic_lattice((256,256)), 64.0, (0.5,0,0,0), 200, 50, tile_sites=64)
