"""Multi-rank (imaginary-time slab) engine on ONE GPU through the in-process loopback
communicator: P engines, P threads.  The collectives are the ones the real multi-GPU run makes
(torch.distributed/NCCL adapter in alps-looper_b200/comm.py); only the transport differs."""
import importlib.util
import os

import numpy as np
import pytest

import oracle_util as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mods():
    import looper_b200 as lq
    spec = importlib.util.spec_from_file_location("lq_comm", os.path.join(ROOT, "alps-looper_b200", "comm.py"))
    comm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(comm)
    return lq, comm


SUMS = ["umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
        "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag"]


@pytest.mark.parametrize("P", [2, 3, 4])
@pytest.mark.parametrize("case", ["chain16", "square8"])
def test_slab_merge_equals_serial_partition(P, case):
    """same configuration loaded into P slab engines: the merged cluster count and all collector
    sums equal the oracle's (reference union-find on the whole configuration)."""
    lq, comm = _mods()
    lat, beta = (lq.chain_lattice(16), 10.0) if case == "chain16" else (lq.hypercubic_lattice((8, 8)), 6.0)
    sim = orc.OracleSim(lat, beta)
    for _ in range(200):
        sim.sweep()
    spins, ops = sim.get_state()
    ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    grp = comm.LoopbackGroup(P)

    def body(r):
        eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=99)
        grp.attach(eng, r)
        eng.set_state(spins, ops)
        nloc = eng.num_ops()
        import ctypes as C
        nc = C.c_int64(0)
        c = lq.LqCollector()
        lq._check(lq.lib.lq_build_clusters(eng._h, None, C.byref(nc), C.byref(c)))
        d = c.as_dict()
        eng.close()
        return nloc, nc.value, d

    res = grp.run(body)
    assert sum(r[0] for r in res) == len(ops)
    for nloc, nc, d in res:
        assert nc == ref_nc
        assert d["nop"] == len(ops)
        for f in SUMS:
            assert d[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f


def test_slab_sweeps_are_legal_and_physical():
    """P=2 slabs sweeping a chain: the union of the slabs stays a legal configuration, every rank
    reports the same collector, and the observables agree with exact diagonalisation."""
    lq, comm = _mods()
    L, T = 8, 0.2
    beta = 1 / T
    lat = lq.chain_lattice(L)
    P = 2
    grp = comm.LoopbackGroup(P)
    nsweeps = 6000

    def body(r):
        eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=4242)
        grp.attach(eng, r)
        eng.sweep_many(500, collect=False)
        out = eng.sweep_many(nsweeps)
        spins, ops = eng.get_state()
        eng.close()
        return out, spins, ops

    res = grp.run(body)
    for f in res[0][0].dtype.names:
        assert np.array_equal(res[0][0][f], res[1][0][f]), f
    ops = np.concatenate([res[0][2], res[1][2]])
    assert np.all(np.diff(ops["time"]) >= 0)
    orc.build_clusters(lat, res[0][1], ops)      # rank 0 holds the spins at tau = 0
    out = res[0][0]
    assert out["nop"][-1] == len(ops)
    ene = (0.25 * L - out["nop"] / beta) / L
    ssus = beta * out["usize"] / L
    smag = out["usize2"]

    def berr(x, nb=30):
        m = len(x) // nb
        b = x[: m * nb].reshape(nb, m).mean(axis=1)
        return b.std(ddof=1) / np.sqrt(nb)

    for name, series, ex in [("energy", ene, -0.441438), ("smag", smag, 6.59939), ("ssus", ssus, 2.40159)]:
        assert abs(series.mean() - ex) < 4 * berr(series) + 1e-12, (name, series.mean(), ex, berr(series))


@pytest.mark.parametrize("P", [2, 3])
def test_slab_merge_with_site_graphs_and_winding_numbers(P):
    """transverse-field + cross-graph configuration from the oracle's generic sweep, loaded into P
    slabs: clusters, susceptibility sums, transmag length (transmag.h) and stiffness sum (stiffness.h)
    of the merged result equal the whole-configuration values -- the open-cluster table of the
    exchange carries the site-leg count and the windings."""
    lq, comm = _mods()
    lat = lq.hypercubic_lattice((6, 6))
    beta = 4.0
    v, off, sign = lq.xxz_weights(-1.0, 0.5)
    sim = orc.OracleModelSim(lat, beta, weights=tuple(v), site_weight=0.3, seed=21)
    for _ in range(150):
        sim.sweep()
    spins, ops = sim.get_state()
    assert ((ops["loc"] & 1) == 0).any()
    ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    ref_w2, _ = orc.stiffness(lat, spins, ops)
    assert ref["tlen"] > 0
    grp = comm.LoopbackGroup(P)

    def body(r):
        eng = lq.Engine(lat, beta, weights=tuple(v), site_weight=0.3, rank=r, nranks=P, seed=99, stiffness=True)
        grp.attach(eng, r)
        eng.set_state(spins, ops)
        import ctypes as C
        nc = C.c_int64(0)
        c = lq.LqCollector()
        lq._check(lq.lib.lq_build_clusters(eng._h, None, C.byref(nc), C.byref(c)))
        d = c.as_dict()
        # a few steps of the slab engine itself keep the union of the slabs legal
        eng.sweep_many(10, collect=False)
        s2, o2 = eng.get_state()
        eng.close()
        return nc.value, d, s2, o2

    res = grp.run(body)
    for nc, d, _, _ in res:
        assert nc == ref_nc
        for f in SUMS + ["tlen"]:
            assert d[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f
        assert d["w2"] == pytest.approx(ref_w2, rel=1e-10, abs=1e-10)
    allops = np.concatenate([r[3] for r in res])
    allops = allops[np.argsort(allops["time"], kind="stable")]
    orc.build_clusters(lat, res[0][2], allops)     # raises on an illegal configuration


def test_slab_engines_rewind_together_and_follow_beta():
    """Slab engines are no second-class citizens (VERDICT r01 missing 4): an arena that overflows on
    one rank makes ALL ranks rewind, grow and replay (the flag travels with the boundary ids, before
    anything is flipped) -- the Markov chain equals the one of amply sized engines; lq_set_beta
    re-buckets every slab in place (the slab boundaries r / P do not move with beta)."""
    lq, comm = _mods()
    lat = lq.hypercubic_lattice((12, 12))
    beta, P = 6.0, 2

    def run(reserve, cluster_reserve):
        grp = comm.LoopbackGroup(P)

        def body(r):
            eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=99, tile_sites=16, reserve=reserve,
                            cluster_reserve=cluster_reserve)
            grp.attach(eng, r)
            out = eng.sweep_many(25)
            regrows = eng.regrow_count()
            s1, o1 = eng.get_state()
            eng.set_beta(9.0)
            s2, o2 = eng.get_state()
            assert np.array_equal(s1, s2) and np.array_equal(o1, o2)
            out2 = eng.sweep_many(30)
            s3, o3 = eng.get_state()
            eng.close()
            return out, regrows, (s1, o1), out2, (s3, o3)

        return grp.run(body)

    small, ample = run(0.25, 0.05), run(0.0, 0.0)
    assert max(r[1] for r in small) > 0 and max(r[1] for r in ample) == 0
    for r in range(P):
        for f in ("nop", "nc", "noc"):
            assert np.array_equal(small[r][0][f], ample[r][0][f]), f
        assert np.array_equal(small[r][2][0], ample[r][2][0]) and np.array_equal(small[r][2][1], ample[r][2][1])
    # after set_beta: the union of the slabs is a legal configuration with more operators
    ops = np.concatenate([ample[r][4][1] for r in range(P)])
    assert np.all(np.diff(ops["time"]) >= 0)
    orc.build_clusters(lat, ample[0][4][0], ops)
    assert ample[0][3]["nop"][-10:].mean() > 1.2 * ample[0][0]["nop"][-10:].mean()
