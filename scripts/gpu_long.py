"""Long GPU run of a small lattice: means with blocked errors (for comparison with the oracle)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import looper_b200 as lq
def berr(x, nb=64):
    m = len(x) // nb; b = np.asarray(x[:m * nb]).reshape(nb, m).mean(axis=1); return b.std(ddof=1) / np.sqrt(nb)
dims = eval(sys.argv[1]); beta = float(sys.argv[2]); n = int(sys.argv[3]); seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tile = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lat = lq.hypercubic_lattice(dims); N = lat["num_sites"]; B = len(lat["src"])
eng = lq.Engine(lat, beta, seed=seed, tile_sites=tile)
eng.sweep_many(2000, collect=False)
out = eng.sweep_many(n)
for name, x in (("usus", beta * out["umag2"] / N), ("smag", out["usize2"]), ("ssus", beta * out["usize"] / N),
                ("ene", (0.25 * B - out["nop"] / beta) / N), ("nc", out["nc"])):
    print("gpu", seed, dims, beta, "tile", tile, name, x.mean(), berr(x), flush=True)
if len(sys.argv) > 6:   # chunk means of the uniform susceptibility (calibration of the error bars)
    x = beta * out["umag2"] / N
    ch = x[: (len(x) // 12000) * 12000].reshape(-1, 12000)
    print("chunk means", np.round(ch.mean(axis=1), 5), "std", ch.mean(axis=1).std(ddof=1), flush=True)
