"""Spin stiffness (looper/stiffness.h:82-133): the per-cluster winding numbers accumulated on the GPU
against the oracle's restatement on the same configurations, and the improved estimator against the
normal one (:137-170) over a Markov chain."""
import json
import os

import numpy as np
import pytest

import oracle_util as orc

pytestmark = pytest.mark.gpu


def _berr(x, nb=32):
    m = len(x) // nb
    b = np.asarray(x[: m * nb]).reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


@pytest.mark.parametrize("jxy,jz,dims", [(1.0, 1.0, (6, 6)), (1.0, 0.5, (6,)), (1.0, 2.0, (4, 4, 4)), (-1.0, -1.0, (6, 6))])
def test_winding_sums_match_oracle(jxy, jz, dims):
    import looper_b200 as lq
    lat = lq.hypercubic_lattice(dims) if len(dims) > 1 else lq.chain_lattice(dims[0])
    v, off, sign = lq.xxz_weights(jxy, jz)
    eng = lq.Engine(lat, 4.0, weights=tuple(v), seed=5, tile_sites=16, stiffness=True)
    nonzero = 0
    for rep in range(8):
        eng.sweep_many(15, collect=False)
        spins, ops = eng.get_state()
        ref, _ = orc.stiffness(lat, spins, ops)
        labels, nc, coll = eng.build_clusters()
        assert coll["w2"] == pytest.approx(ref, rel=1e-12, abs=1e-12)
        nonzero += ref > 0
    assert nonzero > 0
    eng.close()


def test_improved_equals_normal_estimator_on_average():
    import looper_b200 as lq
    lat = lq.hypercubic_lattice((4, 4))
    beta = 2.0
    eng = lq.Engine(lat, beta, seed=23, stiffness=True)
    eng.sweep_many(500, collect=False)
    imp, nrm = [], []
    for i in range(6000):
        c = eng.sweep()                 # improved estimator of this step's clusters
        spins, ops = eng.get_state()
        imp.append(c["w2"])
        nrm.append(orc.stiffness(lat, spins, ops)[1])   # normal estimator of the configuration
    err = np.hypot(_berr(imp), _berr(nrm))
    assert np.mean(imp) > 0.003
    assert abs(np.mean(imp) - np.mean(nrm)) < 4.5 * err, (np.mean(imp), np.mean(nrm), err)
    assert lq.stiffness({"w2": np.mean(imp)}, beta, eng.vector_dim) > 0
    eng.close()


@pytest.mark.parametrize("row", [0, 1, 3])
def test_stiffness_vs_exact_diagonalisation(row):
    """The improved estimator of the engine against tests/golden/ed_stiffness.json (<W^2> = beta F''(0),
    second-order perturbation theory in the twist; RNG-free): Heisenberg ring, XXZ ring with cross graphs,
    4 x 2 ladder.  3 sigma, blocked errors, fixed seeds; the same golden pins the oracle
    (tests/test_oracle_model.py::test_oracle_stiffness_vs_exact_diagonalisation)."""
    import looper_b200 as lq
    ed = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ed_stiffness.json")))[row]
    lat = lq.hypercubic_lattice((4, 2)) if ed["dim"] == 2 else lq.chain_lattice(ed["n"])
    assert [[int(a), int(b)] for a, b in zip(lat["src"], lat["dst"])] == ed["bonds"]
    assert np.allclose(lat["bond_vectors"], ed["rvec"])
    v, off, sign = lq.xxz_weights(ed["jxy"], ed["jz"])
    beta = 1 / ed["T"]
    eng = lq.Engine(lat, beta, weights=tuple(v), seed=41 + row, stiffness=True)
    assert eng.vector_dim == ed["dim"]
    eng.sweep_many(1000, collect=False)
    w2 = eng.sweep_many(60000)["w2"]
    eng.close()
    assert abs(w2.mean() - ed["w2"]) < 3 * _berr(w2), (w2.mean(), ed["w2"], _berr(w2))
    assert lq.stiffness({"w2": w2.mean()}, beta, ed["dim"]) == pytest.approx(ed["stiffness"], abs=3 * _berr(w2) / (beta * ed["dim"]))


def test_without_vectors_the_collector_field_is_zero():
    import looper_b200 as lq
    eng = lq.Engine(lq.chain_lattice(8), 2.0, seed=1)
    out = eng.sweep_many(20)
    assert np.all(out["w2"] == 0)
    eng.close()
