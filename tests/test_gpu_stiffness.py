"""Spin stiffness (looper/stiffness.h:82-133): the per-cluster winding numbers accumulated on the GPU
against the oracle's restatement on the same configurations, and the improved estimator against the
normal one (:137-170) over a Markov chain."""
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


def test_without_vectors_the_collector_field_is_zero():
    import looper_b200 as lq
    eng = lq.Engine(lq.chain_lattice(8), 2.0, seed=1)
    out = eng.sweep_many(20)
    assert np.all(out["w2"] == 0)
    eng.close()
