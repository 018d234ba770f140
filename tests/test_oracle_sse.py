"""CPU tests of the oracle's SSE restatement (oracle.cpp orc_sse_sweep = sse.C:168-407) against
exact diagonalisation (SURVEY Appendix B = loop.op 'diagonalization' blocks; tests/golden/ed_chain.json)
and against the reference's own SSE run in loop.op (lines 331-470: chain L = 4, T = 0.25,
ALGORITHM = "loop; sse"), whose printed values must be statistically compatible with ours.
The estimators are the SSE forms of susceptibility.h:213-215."""
import json
import os

import numpy as np
import pytest

import oracle_util as orc
import looper_lattices as ll

HERE = os.path.dirname(os.path.abspath(__file__))


def _berr(x, nb=32):
    x = np.asarray(x, dtype=np.float64)
    m = len(x) // nb
    b = x[: m * nb].reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


def sse_series(c, beta, vol):
    """susceptibility.h:213-215,247-253 and energy.h:79 from a list of SSE collectors"""
    nop = np.array([x.nop for x in c])
    dip = lambda x: np.where(nop > 0, x / np.maximum(nop, 1), 0.0)
    f = lambda a, a2: beta * (dip(np.array([getattr(x, a) for x in c])) + np.array([getattr(x, a2) for x in c])) / (nop + 1) / vol
    return {
        "energy_density": np.array([x.ene for x in c]) / vol,
        "usus_density": f("umag", "umag2"),
        "smag2": np.array([x.smag2 for x in c]),
        "ssus_density": f("smag", "smag2"),
    }


# loop.op:353,367-378,... the reference's own SSE numbers for the L = 4 chain at T = 0.25 (value, error)
LOOP_OP_SSE_L4 = {"energy_density": (-0.49086, 0.00504), "ssus_density": (1.23722, 0.0112), "smag2": (2.63672, 0.0189)}
ED_L4 = {"energy_density": -0.485876, "usus_density": 0.0359724, "smag2": 2.62731, "ssus_density": 1.25365}


def test_sse_oracle_chain_L4_vs_ed_and_loop_op():
    lat = ll.chain_lattice(4)
    beta = 4.0
    sim = orc.OracleModelSim(lat, beta, weights=(0.5, 0, 0, 0), seed=23125)
    for _ in range(4000):
        sim.sse_sweep()
    c = [sim.sse_sweep() for _ in range(60000)]
    ser = sse_series(c, beta, 4.0)
    for k, ex in ED_L4.items():
        err = _berr(ser[k])
        assert abs(ser[k].mean() - ex) < 4.0 * err + 1e-12, (k, ser[k].mean(), ex, err)
    # the reference's own run (8192 sweeps) agrees with exact diagonalisation -- and hence with us --
    # within its printed errors
    for k, (v, e) in LOOP_OP_SSE_L4.items():
        assert abs(v - ED_L4[k]) < 3.0 * e, (k, v, ED_L4[k], e)


@pytest.mark.parametrize("row", [0, 1, 2])
def test_sse_oracle_xxz_chain_vs_ed(row):
    ed = json.load(open(os.path.join(HERE, "golden", "ed_chain.json")))[row]
    L, beta = ed["L"], 1 / ed["T"]
    v, off, sign = orc.xxz_weights(ed["jxy"], ed["jz"])
    sim = orc.OracleModelSim(ll.chain_lattice(L), beta, weights=tuple(v), seed=99 + row)
    for _ in range(3000):
        sim.sse_sweep()
    c = [sim.sse_sweep() for _ in range(30000)]
    ser = sse_series(c, beta, float(L))
    ser["usus_density"] = beta * np.array([x.umag2 for x in c]) / L   # conserved magnetisation: beta <M^2> / N
    for k in ("energy_density", "usus_density", "smag2", "ssus_density"):
        err = _berr(ser[k])
        assert abs(ser[k].mean() - ed[k]) < 4.0 * err + 1e-12, (k, ser[k].mean(), ed[k], err)


def test_sse_collect_is_the_rescaled_path_integral_collector():
    """cluster sums are linear in the operator times: the SSE collector of a string equals the
    path-integral collector of the same operators at times k/n, times n (n^2 for squared sums)."""
    lat = ll.hypercubic_lattice((4, 4))
    sim = orc.OracleModelSim(lat, 3.0, weights=(0.5, 0, 0, 0), seed=5)
    for _ in range(200):
        sim.sse_sweep()
    spins, ops = sim.get_state()
    n = len(ops)
    assert n > 10 and np.all(np.diff(ops["time"]) > 0)
    a = orc.sse_collect(lat, spins, ops)
    ops2 = ops.copy()
    ops2["time"] = np.arange(n) / n
    _, _, b = orc.build_clusters(lat, spins, ops2)
    assert a["nop"] == n and a["nc"] == b["nc"]
    for f in ("usize", "umag", "ssize", "smag"):
        assert a[f] == pytest.approx(n * n * b[f], rel=1e-12)
    for f in ("usize2", "umag2", "smag2", "umag0", "smag0"):
        assert a[f] == pytest.approx(b[f], rel=1e-12, abs=1e-12)
