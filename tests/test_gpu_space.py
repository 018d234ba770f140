"""Multi-rank engine with the SPATIAL cut (lq_options.cut = LQ_CUT_SPACE; BASELINE config 3 "space-
partitioned", reference analogue looper/lattice.h:692-787) on ONE GPU through the in-process loopback
communicator: P engines, P threads, each owning a contiguous range of tiles over the whole
imaginary-time axis plus ghost copies of the neighbouring tiles.  The collectives and the halo
exchange are the ones the real multi-GPU run makes; only the transport differs."""
import ctypes as C
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

CASES = {
    "chain16": (lambda lq: lq.chain_lattice(16), 10.0, 2),
    "square8": (lambda lq: lq.hypercubic_lattice((8, 8)), 6.0, 4),
    "square12x8": (lambda lq: lq.hypercubic_lattice((12, 8)), 4.0, 8),
    "cubic4": (lambda lq: lq.hypercubic_lattice((4, 4, 4)), 3.0, 8),
}


def _merge_states(lat, res):
    """union of the ranks' owned operators (time-sorted) and own spins"""
    spins = np.full(lat["num_sites"], -1, dtype=np.int32)
    for s, _ in res:
        own = s >= 0
        assert np.all(spins[own] < 0), "a site is owned by two ranks"
        spins[own] = s[own]
    assert np.all(spins >= 0), "a site is owned by no rank"
    ops = np.concatenate([o for _, o in res])
    ops = ops[np.argsort(ops["time"], kind="stable")]
    return spins, ops


@pytest.mark.parametrize("P", [2, 3, 4])
@pytest.mark.parametrize("case", list(CASES))
def test_space_merge_equals_serial_partition(P, case):
    """same configuration loaded into P spatial engines: the merged cluster count and all collector
    sums equal the oracle's (reference union-find on the whole configuration)."""
    lq, comm = _mods()
    mk, beta, ts = CASES[case]
    lat = mk(lq)
    sim = orc.OracleSim(lat, beta)
    for _ in range(200):
        sim.sweep()
    spins, ops = sim.get_state()
    ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    grp = comm.LoopbackGroup(P)

    def body(r):
        eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=99, cut="space", tile_sites=ts)
        grp.attach(eng, r)
        eng.set_state(spins, ops)
        nc = C.c_int64(0)
        c = lq.LqCollector()
        lq._check(lq.lib.lq_build_clusters(eng._h, None, C.byref(nc), C.byref(c)))
        d = c.as_dict()
        st = eng.get_state()
        eng.close()
        return nc.value, d, st

    res = grp.run(body)
    s2, o2 = _merge_states(lat, [r[2] for r in res])
    assert np.array_equal(s2, spins)
    assert len(o2) == len(ops) and np.array_equal(o2["time"], ops["time"]) and np.array_equal(o2["type"], ops["type"])
    for nc, d, _ in res:
        assert nc == ref_nc
        assert d["nop"] == len(ops)
        for f in SUMS:
            assert d[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f


@pytest.mark.parametrize("P,case", [(2, "chain8"), (4, "square8")])
def test_space_sweeps_are_legal_and_physical(P, case):
    """spatial engines sweeping: the union of the ranks' operators stays a legal configuration, every
    rank reports the same collector, and the observables agree with exact diagonalisation (chain) /
    the reference CPU algorithm (square lattice, 3 sigma)."""
    lq, comm = _mods()
    if case == "chain8":
        L, T = 8, 0.2
        beta = 1 / T
        lat = lq.chain_lattice(L)
        ts, nsweeps = 2, 6000
    else:
        lat = lq.hypercubic_lattice((8, 8))
        beta, ts, nsweeps = 2.0, 4, 4000
    grp = comm.LoopbackGroup(P)

    def body(r):
        eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=4242, cut="space", tile_sites=ts)
        grp.attach(eng, r)
        eng.sweep_many(500, collect=False)
        out = eng.sweep_many(nsweeps)
        st = eng.get_state()
        eng.close()
        return out, st

    res = grp.run(body)
    for r in range(1, P):
        for f in res[0][0].dtype.names:
            assert np.array_equal(res[0][0][f], res[r][0][f]), f
    spins, ops = _merge_states(lat, [r[1] for r in res])
    orc.build_clusters(lat, spins, ops)      # raises on an illegal configuration
    out = res[0][0]
    assert out["nop"][-1] == len(ops)
    N = lat["num_sites"]
    nb = len(lat["src"])

    def berr(x, nbk=30):
        m = len(x) // nbk
        b = x[: m * nbk].reshape(nbk, m).mean(axis=1)
        return b.std(ddof=1) / np.sqrt(nbk)

    ene = (0.25 * nb - out["nop"] / beta) / N
    ssus = beta * out["usize"] / N
    smag = out["usize2"]
    if case == "chain8":
        exact = [("energy", ene, -0.441438), ("smag", smag, 6.59939), ("ssus", ssus, 2.40159)]
        for name, series, ex in exact:
            assert abs(series.mean() - ex) < 4 * berr(series) + 1e-12, (name, series.mean(), ex, berr(series))
    else:
        sim = orc.OracleSim(lat, beta, seed=7)
        for _ in range(500):
            sim.sweep()
        c = [sim.sweep() for _ in range(nsweeps)]
        cpu = {f: np.array([x[f] for x in c]) for f in ("nop", "nc", "sa_usus", "sa_smag", "sa_ssus")}
        # (looper-named sums vs standalone/loop.C:173-178: umag2 = usus/4, usize2 = smag/4, usize = ssus/4)
        pairs = {"nop": (out["nop"], cpu["nop"]), "clusters": (out["nc"], cpu["nc"]),
                 "uniform susceptibility": (out["umag2"], 0.25 * cpu["sa_usus"]),
                 "staggered magnetization^2": (out["usize2"], 0.25 * cpu["sa_smag"]),
                 "staggered susceptibility": (out["usize"], 0.25 * cpu["sa_ssus"])}
        for k, (a, b) in pairs.items():
            err = np.hypot(berr(a), berr(b))
            assert abs(a.mean() - b.mean()) < 3.0 * err + 1e-12, (k, a.mean(), b.mean(), err)


@pytest.mark.parametrize("P", [2, 3])
def test_space_merge_with_site_graphs_and_winding_numbers(P):
    """transverse-field + cross-graph configuration from the oracle's generic sweep, loaded into P
    spatial engines: clusters, susceptibility sums, transmag length and stiffness sum of the merged
    result equal the whole-configuration values; a few steps keep the union legal."""
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
    grp = comm.LoopbackGroup(P)

    def body(r):
        eng = lq.Engine(lat, beta, weights=tuple(v), site_weight=0.3, rank=r, nranks=P, seed=99, stiffness=True,
                        cut="space", tile_sites=4)
        grp.attach(eng, r)
        eng.set_state(spins, ops)
        nc = C.c_int64(0)
        c = lq.LqCollector()
        lq._check(lq.lib.lq_build_clusters(eng._h, None, C.byref(nc), C.byref(c)))
        d = c.as_dict()
        eng.sweep_many(10, collect=False)
        st = eng.get_state()
        eng.close()
        return nc.value, d, st

    res = grp.run(body)
    for nc, d, _ in res:
        assert nc == ref_nc
        for f in SUMS + ["tlen"]:
            assert d[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f
        assert d["w2"] == pytest.approx(ref_w2, rel=1e-10, abs=1e-10)
    s2, o2 = _merge_states(lat, [r[2] for r in res])
    orc.build_clusters(lat, s2, o2)     # raises on an illegal configuration


def test_space_engines_rewind_together_and_follow_beta():
    """an arena that overflows on one rank makes ALL ranks rewind, grow and replay -- the Markov chain
    equals the one of amply sized engines; lq_set_beta re-buckets every rank's tiles in place."""
    lq, comm = _mods()
    lat = lq.hypercubic_lattice((12, 12))
    beta, P = 6.0, 2

    def run(reserve, cluster_reserve):
        grp = comm.LoopbackGroup(P)

        def body(r):
            eng = lq.Engine(lat, beta, rank=r, nranks=P, seed=99, tile_sites=16, reserve=reserve,
                            cluster_reserve=cluster_reserve, cut="space")
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
    spins, ops = _merge_states(lat, [ample[r][4] for r in range(P)])
    orc.build_clusters(lat, spins, ops)
    assert ample[0][3]["nop"][-10:].mean() > 1.2 * ample[0][0]["nop"][-10:].mean()
