"""XXZ bonds (graph types 0..3 of graph_impl.h:255-328: horizontal, cross, frozen) on the GPU:
partition parity with the oracle's reconnect rules on configurations the GPU produced, and
observables against exact diagonalisation (tests/golden/ed_chain.json, made by
tests/golden/make_ed_golden.py -- a numpy restatement of diag.C:376-468)."""
import json
import os

import numpy as np
import pytest

import oracle_util as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SUMS = ["umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
        "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag"]


@pytest.mark.parametrize("jz,lat_kind", [(0.5, "chain"), (2.0, "chain"), (0.0, "square"), (1.5, "square")])
def test_xxz_partition_matches_reference_reconnect(jz, lat_kind):
    import looper_b200 as lq
    lat = lq.chain_lattice(16) if lat_kind == "chain" else lq.hypercubic_lattice((6, 6))
    v, off, sign = lq.xxz_weights(1.0, jz)
    eng = lq.Engine(lat, 6.0, weights=tuple(v), seed=77, tile_sites=16)
    assert eng.info()["nodes_per_op"] == (2 if v[1] > 0 else 1)
    seen = set()
    for rep in range(6):
        eng.sweep_many(25, collect=False)
        spins, ops = eng.get_state()
        seen |= set((ops["type"] >> 2).tolist())
        ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)   # raises on an illegal string
        labels, nc, coll = eng.build_clusters()
        assert nc == ref_nc
        assert np.array_equal(labels, ref_labels)
        for f in SUMS:
            assert coll[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f
    expect = {0, 1} if abs(jz) < 1 else {0, 2}
    assert seen == expect, seen
    eng.close()


def _berr(x, nb=32):
    m = len(x) // nb
    b = np.asarray(x[: m * nb]).reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


@pytest.mark.parametrize("row", [1, 2, 3, 4])
def test_xxz_observables_vs_exact_diagonalisation(row):
    import looper_b200 as lq
    ed = json.load(open(os.path.join(HERE, "golden", "ed_chain.json")))[row]
    L, T = ed["L"], ed["T"]
    beta = 1 / T
    v, off, sign = lq.xxz_weights(ed["jxy"], ed["jz"])
    lat = lq.chain_lattice(L)
    eng = lq.Engine(lat, beta, weights=tuple(v), seed=2024 + row)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(24000)
    eng.close()
    series = {
        "energy_density": out["ene"] / L,                 # energy.h:79
        "usus_density": beta * out["umag2"] / L,          # conserved M: beta <M^2> / N
        "smag2": out["smag2"],                            # susceptibility.h:241
        "ssus_density": beta * out["smag"] / L,           # susceptibility.h:247
    }
    for k, x in series.items():
        err = _berr(x)
        assert abs(x.mean() - ed[k]) < 4.5 * err + 1e-10, (k, x.mean(), ed[k], err)
