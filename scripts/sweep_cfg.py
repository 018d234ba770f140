#!/usr/bin/env python
"""Phase times per operator for a list of (tile_sites, window_ops, reserve) engine configurations
on one mid-size workload (GPU box only).  usage: sweep_cfg.py L beta therm "tile,wops,reserve" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import looper_b200 as lq  # noqa: E402

L, beta, therm = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
lat = lq.hypercubic_lattice((L, L))
for spec in sys.argv[4:]:
    tile, wops, res = spec.split(",")
    try:
        eng = lq.Engine(lat, beta, seed=29833, tile_sites=int(tile), window_ops=float(wops), reserve=float(res))
        eng.sweep_many(therm, collect=False)
        if os.environ.get("LQ_DBG"):
            import ctypes as C
            buf = (C.c_uint64 * 8)()
            lq.lib.lq_debug_counters(eng._h, buf)   # clear
        eng.enable_timers(True)
        out = eng.sweep_many(6)
        nop = float(out["nop"].mean())
        tm = {t["id"]: 1e12 * t["seconds"] / t["count"] / nop for t in eng.timers()}
        info = eng.info()
        tot = sum(tm.values())
        print(f"cfg tile={tile} wops={wops} res={res} W={info['num_windows']} cap={info['page_capacity']} "
              f"tpb={info['threads_per_page']} GB={info['device_bytes'] / 1e9:.1f} nop={nop:.3g} "
              f"ps/op: total={tot:.1f} " + " ".join(f"{k}:{v:.1f}" for k, v in sorted(tm.items())), flush=True)
        if os.environ.get("LQ_DBG"):
            lq.lib.lq_debug_counters(eng._h, buf)
            c = list(buf)
            print(f"  dbg: global edges/op={c[0] / 6 / nop:.3f} hops/edge={c[1] / max(c[0], 1):.2f} "
                  f"retries/edge={c[2] / max(c[0], 1):.4f} edges>16hops={c[3] / max(c[0], 1):.4f} "
                  f"compress hops/node={c[5] / max(c[4], 1):.2f}", flush=True)
        eng.close()
    except Exception as e:  # noqa: BLE001
        print(f"cfg {spec} FAILED: {e}", flush=True)
