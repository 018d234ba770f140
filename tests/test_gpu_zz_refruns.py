"""The GPU engine against the reference's OWN Monte Carlo results (tests/golden/ref_runs.json, extracted from
loop.op and extras/*/*.op by tests/golden/make_ref_goldens.py; the oracle is pinned to the same numbers in
tests/test_oracle_refruns.py): "Number of Clusters", energy, magnetisations, susceptibilities, "Stiffness"
and "Transverse Magnetization" of the path-integral tasks on chains and single sites, 4 sigma of the combined
error (the reference ran 1024 or 4096 sweeps; its error dominates)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS = json.load(open(os.path.join(HERE, "golden", "ref_runs.json")))
# Reference runs that sit far from the exact value themselves (so that "4 sigma from the reference" would be a coin
# toss): the transverse-field Ising chain of extras/gap (one Markov chain, printed twice: staggered magnetisation^2
# 1.567 +- 0.026 against 1.6567 by exact diagonalisation, 3.5 sigma) and the Ising chain of loop.op (cluster count
# 4 sigma from its own SSE twin).  They are compared at 6 sigma -- still a factor-of-two test of every observable.
NSIGMA = {"extras/gap/gap.op:619": 6.0, "extras/gap/gap.op:1257": 6.0, "loop.op:2070": 6.0}
PICK = [i for i, r in enumerate(RUNS) if r["algorithm"] == "loop; path integral" and r["improved"]
        and r["lattice"] == "chain lattice" and r["source"].split("/")[-1].split(":")[0] in ("loop.op", "transmag.op", "gap.op")
        # (loop.op:2070, the Ising chain: the reference's "Number of Clusters" of that run, 2.092 +- 0.014, is 4 sigma from its
        # own SSE twin loop.op:2324, 2.164 +- 0.018 -- an underestimated error bar; the oracle test keeps the case, here it
        # would be a coin toss at 4 sigma)
        and r["source"] != "loop.op:2070"]


def _berr(x, nb=32):
    m = len(x) // nb
    b = np.asarray(x[: m * nb]).reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


@pytest.mark.parametrize("i", PICK, ids=[RUNS[i]["source"] for i in PICK])
def test_engine_against_the_reference_run(i):
    import looper_b200 as lq
    r = RUNS[i]
    lat = lq.chain_lattice(r["L"])
    vol, beta = lat["num_sites"], 1 / r["T"]
    want_stiff = "Stiffness" in r["results"]
    eng = lq.Engine(lat, beta, weights=tuple(lq.xxz_weights(r["Jxy"], r["Jz"])[0]), site_weight=r["Gamma"] / 2,
                    seed=300 + i, stiffness=want_stiff)
    eng.sweep_many(1000, collect=False)
    out = eng.sweep_many(12000)
    eng.close()
    series = {}
    for c in out:
        o = lq.observables(c, beta, vol)
        if want_stiff:
            o["Stiffness"] = lq.stiffness(c, beta, 1)
        for k in r["results"]:
            if k in o:
                series.setdefault(k, []).append(o[k])
    assert "Number of Clusters" in series and "Energy" in series
    for k, x in series.items():
        g = r["results"][k]
        err = np.hypot(g["error"], _berr(x))
        assert abs(np.mean(x) - g["value"]) < NSIGMA.get(r["source"], 4.0) * err + 1e-12, (r["source"], k, np.mean(x), g, _berr(x))
