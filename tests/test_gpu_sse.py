"""The SSE representation on the GPU (lq_options.representation = LQ_REPR_SSE; reference: sse.C:168-407,
susceptibility.h:213-215).  An SSE string is the time-ordered operator list of a world-line
configuration; the engine produces the string positions with a counting sort (csrc/lq_kernels.cuh
k_sse_*) and accumulates the fixed-length-string estimators with integer times.

* cluster sums of the GPU's own strings against the oracle's SSE collector -- integer arithmetic on
  both sides of the comparison, so the tolerance is f64 rounding only;
* observables against exact diagonalisation (chain, 4 x 2 ladder) and, for the cubic lattice, against
  the oracle's restatement of the SSE worker (orc_sse_sweep) within 3 sigma."""
import json
import os

import numpy as np
import pytest

import oracle_util as orc
from test_oracle_sse import sse_series, _berr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SUMS = ["umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
        "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag"]


class _C:   # collector rows of sweep_many as attribute objects for sse_series
    def __init__(self, row):
        self.__dict__.update({f: float(row[f]) for f in row.dtype.names})


@pytest.mark.parametrize("case", ["chain16", "square12_tiles", "cubic6", "xxz_chain", "tfi_chain"])
def test_sse_collector_matches_oracle_on_gpu_strings(case):
    import looper_b200 as lq
    kw = {}
    if case == "chain16":
        lat, beta = lq.chain_lattice(16), 10.0
    elif case == "square12_tiles":
        lat, beta, kw = lq.hypercubic_lattice((12, 12)), 6.0, dict(tile_sites=16)
    elif case == "cubic6":
        lat, beta = lq.hypercubic_lattice((6, 6, 6)), 1.5
    elif case == "xxz_chain":
        lat, beta = lq.chain_lattice(16), 6.0
        kw = dict(weights=tuple(lq.xxz_weights(1.0, 0.5)[0]), tile_sites=8)
    else:
        lat, beta = lq.chain_lattice(12), 4.0
        kw = dict(weights=tuple(lq.xxz_weights(-1.0, 0.5)[0]), site_weight=0.35)
    eng = lq.Engine(lat, beta, seed=31, sse=True, **kw)
    for rep in range(4):
        out = eng.sweep_many(25)
        spins, ops = eng.get_state()
        ref = orc.sse_collect(lat, spins, ops)
        labels, nc, coll = eng.build_clusters()
        assert nc == ref["nc"] and coll["nop"] == len(ops)
        for f in SUMS + ["tlen"]:
            assert coll[f] == pytest.approx(ref[f], rel=1e-11, abs=1e-9), (f, coll[f], ref[f])
        # the partition is that of the path-integral view of the same configuration
        ref_labels, ref_nc, _ = orc.build_clusters(lat, spins, ops)
        assert nc == ref_nc and np.array_equal(labels, ref_labels)
    eng.close()


def test_sse_observables_vs_exact_diagonalisation_chain():
    import looper_b200 as lq
    ed = json.load(open(os.path.join(HERE, "golden", "ed_chain.json")))[0]      # L = 8, T = 0.2 Heisenberg
    L, beta = ed["L"], 1 / ed["T"]
    eng = lq.Engine(lq.chain_lattice(L), beta, seed=4, sse=True)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(30000)
    eng.close()
    ser = sse_series([_C(r) for r in out], beta, float(L))
    for k in ("energy_density", "usus_density", "smag2", "ssus_density"):
        err = _berr(ser[k])
        assert abs(ser[k].mean() - ed[k]) < 4.0 * err + 1e-12, (k, ser[k].mean(), ed[k], err)
    # the library's own commit formulas (observables(..., sse=True)) give the same numbers
    o = lq.observables({f: float(out[f][-1]) for f in out.dtype.names}, beta, L, sse=True)
    assert o["Staggered Susceptibility"] == pytest.approx(ser["ssus_density"][-1], rel=1e-12)


def test_sse_observables_vs_exact_diagonalisation_ladder():
    import looper_b200 as lq
    ed = json.load(open(os.path.join(HERE, "golden", "ed_ladder.json")))[0]     # 4 x 2 Heisenberg ladder
    lat = lq.hypercubic_lattice((4, 2))
    n, beta = ed["n"], 1 / ed["T"]
    v, off, sign = lq.xxz_weights(ed["jxy"], ed["jz"])
    eng = lq.Engine(lat, beta, weights=tuple(v), site_weight=ed["gamma"] / 2, seed=778, sse=True)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(30000)
    eng.close()
    ser = sse_series([_C(r) for r in out], beta, float(n))
    for k, name in (("energy_density", "energy_density"), ("ssus_density", "ssus_density"), ("smag2", "smag2"),
                    ("usus_density", "usus_density")):
        err = _berr(ser[k])
        assert abs(ser[k].mean() - ed[name]) < 4.0 * err + 1e-10, (k, ser[k].mean(), ed[name], err)


def test_sse_cubic_lattice_agrees_with_the_reference_sse_worker():
    """simple cubic 4 x 4 x 4 near T_N (BASELINE config 4 at reduced size): GPU SSE estimators against
    the oracle's restatement of sse.C within 3 sigma of the combined blocked errors."""
    import looper_b200 as lq
    lat = lq.hypercubic_lattice((4, 4, 4))
    beta, N = 1 / 0.95, 64
    eng = lq.Engine(lat, beta, seed=12345, sse=True)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(40000)
    eng.close()
    g = sse_series([_C(r) for r in out], beta, float(N))
    sim = orc.OracleModelSim(lat, beta, weights=(0.5, 0, 0, 0), seed=777)
    for _ in range(3000):
        sim.sse_sweep()
    c = sse_series([sim.sse_sweep() for _ in range(40000)], beta, float(N))
    for k in ("energy_density", "usus_density", "smag2", "ssus_density"):
        err = np.hypot(_berr(g[k]), _berr(c[k]))
        assert abs(g[k].mean() - c[k].mean()) < 3.0 * err + 1e-12, (k, g[k].mean(), c[k].mean(), err)


def test_sse_needs_a_serial_engine():
    import looper_b200 as lq
    with pytest.raises(lq.LqError) as e:
        lq.Engine(lq.chain_lattice(8), 2.0, sse=True, rank=0, nranks=2)
    assert e.value.code == -6
