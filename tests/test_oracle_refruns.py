"""The generic-model restatement (oracle/oracle.cpp orc_model_*: XXZ bond graphs, frozen graphs, site graphs,
path integral and SSE) against the reference's OWN Monte Carlo results: every S = 1/2 task of the golden
outputs of its regression tests -- loop.op, extras/{transmag,gap,corrlen,top,localsus}/*.op -- extracted by
tests/golden/make_ref_goldens.py into tests/golden/ref_runs.json (value and error as ALPS printed them).
Besides the physics (which exact diagonalisation pins more sharply) these carry what only the reference's
algorithm defines: "Number of Clusters", the generalised magnetisations of the improved estimator, the
winding-number "Stiffness" and the "Transverse Magnetization" of transmag.h.

Statistical goldens of 1024 or 4096 sweeps: the comparison is |oracle - reference| < 4 sigma of the combined
error.  (The reference's own error bars are not always converged -- ALPS says so in the files -- and one of
its runs sits 2.3 sigma from its own exact diagonalisation; that is asserted below as well, so that the
threshold is not mistaken for slack in the restatement.)"""
import importlib.util
import json
import os

import numpy as np
import pytest

import oracle_util as orc
from looper_lattices import chain_lattice, hypercubic_lattice
from oracle_util import xxz_weights

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS = json.load(open(os.path.join(HERE, "golden", "ref_runs.json")))
# Reference runs that sit far from the exact value themselves (so that "4 sigma from the reference" would be a coin
# toss): the transverse-field Ising chain of extras/gap (one Markov chain, printed twice: staggered magnetisation^2
# 1.567 +- 0.026 against 1.6567 by exact diagonalisation, 3.5 sigma) and the Ising chain of loop.op (cluster count
# 4 sigma from its own SSE twin).  They are compared at 6 sigma -- still a factor-of-two test of every observable.
NSIGMA = {"extras/gap/gap.op:619": 6.0, "extras/gap/gap.op:1257": 6.0, "loop.op:2070": 6.0}
QMC = [i for i, r in enumerate(RUNS) if r["algorithm"] != "diagonalization"]


def _berr(x, nb=32):
    m = len(x) // nb
    b = np.asarray(x[: m * nb]).reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


def _observables(c, beta, vol, sse):
    """collector.commit of energy.h:74-83, susceptibility.h:199-254 (path-integral and SSE forms), transmag.h:106-109"""
    nop = c.nop
    dip = (lambda x: x / nop if nop > 0 else 0.0)
    sus = (lambda x, x2: beta * (dip(x) + x2) / (nop + 1) / vol) if sse else (lambda x, x2: beta * x / vol)
    return {"Energy": c.ene, "Energy Density": c.ene / vol, "Number of Clusters": c.nc,
            "Magnetization^2": c.umag2, "Magnetization^4": 3 * c.umag2 ** 2 - 2 * c.umag4,
            "Susceptibility": sus(c.umag, c.umag2),
            "Staggered Magnetization^2": c.smag2, "Staggered Magnetization^4": 3 * c.smag2 ** 2 - 2 * c.smag4,
            "Staggered Susceptibility": sus(c.smag, c.smag2),
            "Generalized Magnetization^2": c.usize2, "Generalized Susceptibility": sus(c.usize, c.usize2),
            "Transverse Magnetization": 0.5 * c.tlen,
            "Energy^2": c.ene ** 2 - nop / beta ** 2}      # energy.h:80


def _lattice(r):
    if r["lattice"] == "site":
        return dict(num_sites=1, src=np.zeros(0, np.int32), dst=np.zeros(0, np.int32), gauge=np.ones(1))
    if r["lattice"] == "chain lattice":
        return chain_lattice(r["L"])
    return hypercubic_lattice((r["L"],) * 3)


@pytest.mark.parametrize("i", QMC, ids=[RUNS[i]["source"] for i in QMC])
def test_oracle_against_the_reference_run(i):
    r = RUNS[i]
    sse = r["algorithm"] == "loop; sse"
    lat = _lattice(r)
    vol, beta = lat["num_sites"], 1 / r["T"]
    v = xxz_weights(r["Jxy"], r["Jz"])[0] if vol > 1 else (0, 0, 0, 0)
    sim = orc.OracleModelSim(lat, beta, weights=tuple(v), site_weight=r["Gamma"] / 2, seed=100 + i)
    nsweeps = 8000 if vol > 16 else 40000
    series = {}
    for s in range(nsweeps):
        c = sim.sse_sweep() if sse else sim.sweep()
        if s < nsweeps // 10:
            continue
        o = _observables(c, beta, vol, sse)
        for k in r["results"]:
            if k in o:
                series.setdefault(k, []).append(o[k])
        if "Stiffness" in r["results"] and not sse:   # stiffness.h:126-129: w2 / (beta dim), improved estimator
            spins, ops = sim.get_state()
            series.setdefault("Stiffness", []).append(orc.stiffness(lat, spins, ops)[0] / (beta * lat["vector_dim"]))
    assert "Number of Clusters" in series and "Energy" in series
    for k, x in series.items():
        g = r["results"][k]
        err = np.hypot(g["error"], _berr(x))
        assert abs(np.mean(x) - g["value"]) < NSIGMA.get(r["source"], 4.0) * err + 1e-12, (r["source"], k, np.mean(x), g, _berr(x))
    # the evaluated observables (energy.h:89-102, susceptibility.h:340-376), jackknife over 32 blocks
    derived = {"Specific Heat": (("Energy", "Energy^2"), lambda e, e2: beta ** 2 * (e2 - e * e) / vol),
               "Binder Ratio of Magnetization": (("Magnetization^2", "Magnetization^4"), lambda a, b: a * a / b),
               "Binder Ratio of Staggered Magnetization": (("Staggered Magnetization^2", "Staggered Magnetization^4"), lambda a, b: a * a / b)}
    for k, (ops, f) in derived.items():
        if k not in r["results"] or any(o not in series for o in ops) or not np.isfinite(r["results"][k]["error"]):
            continue
        blocks = [np.asarray(series[o][: (len(series[o]) // 32) * 32]).reshape(32, -1).mean(axis=1) for o in ops]
        tot = [b.sum() for b in blocks]
        jk = np.array([f(*[(t - b[i]) / 31 for t, b in zip(tot, blocks)]) for i in range(32)])
        val, jerr = f(*[t / 32 for t in tot]), np.sqrt(31 * jk.var())
        g = r["results"][k]
        assert abs(val - g["value"]) < NSIGMA.get(r["source"], 4.0) * np.hypot(g["error"], jerr) + 1e-12, (r["source"], k, val, g, jerr)


def test_numpy_ed_reproduces_the_reference_diagonalization_blocks():
    """tests/golden/make_ed_golden.py (the generator of ed_*.json) against the reference's exact numbers
    (diag.C through LAPACK, loop.op 'diagonalization' tasks): 6 printed digits."""
    spec = importlib.util.spec_from_file_location("make_ed_golden", os.path.join(HERE, "golden", "make_ed_golden.py"))
    ed = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ed)
    seen = 0
    for r in RUNS:
        if r["algorithm"] != "diagonalization" or r["lattice"] != "chain lattice":
            continue
        L = r["L"]
        mine = ed.ed_tfi(L, r["Jxy"], r["Jz"], r["Gamma"], r["T"])
        ref = r["results"]
        assert mine["energy_density"] * L == pytest.approx(ref["Energy"]["value"], rel=5e-6)
        assert mine["umag2"] == pytest.approx(ref["Magnetization^2"]["value"], rel=5e-6)
        assert mine["smag2"] == pytest.approx(ref["Staggered Magnetization^2"]["value"], rel=5e-6)
        assert mine["usus_density"] == pytest.approx(ref["Susceptibility"]["value"], rel=5e-6)       # printed per site
        assert mine["ssus_density"] == pytest.approx(ref["Staggered Susceptibility"]["value"], rel=5e-6)
        seen += 1
    assert seen >= 2


def test_the_reference_runs_scatter_around_its_own_exact_numbers():
    """How sharp these goldens are: the reference's Monte Carlo against its exact diagonalisation of the same
    task (loop.op: Heisenberg chain, Ising chain).  The Ising run sits 2.3 sigma off in the staggered
    magnetisation -- the goldens are samples, not truths."""
    by = {(r["algorithm"], r["L"], r["T"], r["Jz"], r["Jxy"], r["Gamma"]): r for r in RUNS if r["source"].startswith("loop.op")}
    worst = 0.0
    for key, r in by.items():
        if key[0] != "loop; path integral":
            continue
        ex = by.get(("diagonalization",) + key[1:])
        if ex is None:
            continue
        for k in ("Energy", "Magnetization^2", "Staggered Magnetization^2"):
            z = abs(r["results"][k]["value"] - ex["results"][k]["value"]) / r["results"][k]["error"]
            worst = max(worst, z)
            assert z < 4, (r["source"], k, z)
    assert worst > 2
