"""Site graphs (transverse field; graph_impl.h:67-87, path_integral.C:552-561,716-726) and the
transverse magnetisation (transmag.h:62-117) on the GPU: partition and collector parity with the
oracle on configurations from both sides, and observables against exact diagonalisation
(tests/golden/ed_tfi.json, tests/golden/make_ed_golden.py)."""
import json
import os

import numpy as np
import pytest

import oracle_util as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SUMS = ["umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
        "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag", "tlen"]


def _berr(x, nb=32):
    m = len(x) // nb
    b = np.asarray(x[: m * nb]).reshape(nb, m).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(nb)


CASES = [  # (jxy, jz, gamma, lattice)
    (0.0, 1.0, 0.7, "chain"),      # Ising bonds (frozen graph 2) + site graphs: BASELINE config 5 (ii)
    (0.0, -1.0, 0.5, "square"),    # frozen graph 3
    (-1.0, 0.5, 0.6, "chain"),     # horizontal + cross graphs (two nodes per operator) + site graphs
    (-1.0, -1.0, 1.0, "square"),   # cross graphs only
    (0.0, 0.0, 0.9, "chain"),      # free spins in a field: site operators only
]


@pytest.mark.parametrize("jxy,jz,gamma,lat_kind", CASES)
def test_site_graph_partition_and_collector_match_oracle(jxy, jz, gamma, lat_kind):
    import looper_b200 as lq
    lat = lq.chain_lattice(16) if lat_kind == "chain" else lq.hypercubic_lattice((6, 6))
    v, off, sign = lq.xxz_weights(jxy, jz)
    eng = lq.Engine(lat, 5.0, weights=tuple(v), site_weight=gamma / 2, seed=91, tile_sites=16)
    n_site = n_site_off = 0
    for rep in range(6):
        eng.sweep_many(20, collect=False)
        spins, ops = eng.get_state()
        site = (ops["loc"] & 1) == 0
        n_site += int(site.sum())
        n_site_off += int((site & ((ops["type"] & 1) == 1)).sum())
        ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)   # raises on an illegal string
        labels, nc, coll = eng.build_clusters()
        assert nc == ref_nc
        assert np.array_equal(labels, ref_labels)
        for f in SUMS:
            assert coll[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f
        assert coll["nop"] == len(ops)
    assert n_site > 0 and n_site_off > 0
    eng.close()


def test_oracle_configuration_injected_into_gpu():
    import looper_b200 as lq
    lat = lq.hypercubic_lattice((4, 4))
    v, off, sign = lq.xxz_weights(-1.0, 0.5)
    sim = orc.OracleModelSim(lat, 3.0, weights=tuple(v), site_weight=0.4, seed=7)
    eng = lq.Engine(lat, 3.0, weights=tuple(v), site_weight=0.4, seed=1, tile_sites=4)
    for rep in range(5):
        for _ in range(30):
            sim.sweep()
        spins, ops = sim.get_state()
        assert ((ops["loc"] & 1) == 0).any()
        eng.set_state(spins, ops)
        s2, o2 = eng.get_state()
        assert np.array_equal(s2, spins)
        assert np.array_equal(o2["time"], ops["time"]) and np.array_equal(o2["loc"], ops["loc"])
        assert np.array_equal(o2["type"], ops["type"])
        ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)
        labels, nc, coll = eng.build_clusters()
        assert nc == ref_nc and np.array_equal(labels, ref_labels)
        for f in SUMS:
            assert coll[f] == pytest.approx(ref[f], rel=1e-8, abs=1e-7), f
    eng.close()


@pytest.mark.parametrize("row", [0, 1, 2, 3])
def test_tfi_observables_vs_exact_diagonalisation(row):
    import looper_b200 as lq
    ed = json.load(open(os.path.join(HERE, "golden", "ed_tfi.json")))[row]
    L, beta = ed["L"], 1 / ed["T"]
    v, off, sign = lq.xxz_weights(ed["jxy"], ed["jz"])
    assert sign == 1
    eng = lq.Engine(lq.chain_lattice(L), beta, weights=tuple(v), site_weight=ed["gamma"] / 2, seed=4242 + row)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(24000)
    eng.close()
    series = {
        "energy_density": out["ene"] / L,
        "umag2": out["umag2"],
        "smag2": out["smag2"],
        "usus_density": beta * out["umag"] / L,
        "ssus_density": beta * out["smag"] / L,
    }
    if v[2] == 0 and v[3] == 0:   # see tests/test_oracle_model.py on frozen graphs
        series["transmag_density"] = 0.5 * out["tlen"] / L
    for k, x in series.items():
        err = _berr(x)
        assert abs(x.mean() - ed[k]) < 4.5 * err + 1e-10, (k, x.mean(), ed[k], err)


def test_single_spin_in_a_field():
    import looper_b200 as lq
    lat = dict(num_sites=2, src=np.array([0], np.int32), dst=np.array([1], np.int32),
               gauge=np.array([1.0, -1.0]), dims=(2, 0, 0))
    beta, gamma = 1.3, 0.9
    eng = lq.Engine(lat, beta, weights=(0, 0, 0, 0), site_weight=gamma / 2, seed=12)
    eng.sweep_many(500, collect=False)
    out = eng.sweep_many(40000)
    eng.close()
    sx = 0.5 * np.tanh(beta * gamma / 2)
    tm = 0.5 * out["tlen"] / 2
    assert abs(tm.mean() - sx) < 4.5 * _berr(tm)
    en = out["ene"] / 2
    assert abs(en.mean() + gamma * sx) < 4.5 * _berr(en)


def _tfi_chain_energy(jz, gamma, T):
    """Energy per site of the infinite S=1/2 chain H = Jz sum Sz Sz - Gamma sum Sx (free fermions:
    eps_k = 2 sqrt(J^2 + h^2 - 2 J h cos k) with J = Jz/4, h = Gamma/2)."""
    J, h = jz / 4.0, gamma / 2.0
    k = (np.arange(400000) + 0.5) * np.pi / 400000
    eps = 2 * np.sqrt(J * J + h * h - 2 * J * h * np.cos(k))
    return float(-(eps / 2 * np.tanh(eps / (2 * T))).mean())


def test_config5_transverse_field_ising_chain_full_size():
    """BASELINE config 5 (ii) at its full size: chain L = 4096, beta = 256, Jz = 1, Gamma = 0.7.
    Size-independent checks: the energy against the exact free-fermion value, and the partition /
    collector of the million-operator configuration against the oracle."""
    import looper_b200 as lq
    L, beta, gamma = 4096, 256.0, 0.7
    lat = lq.chain_lattice(L)
    v, off, sign = lq.xxz_weights(0.0, 1.0)
    assert v == [0.0, 0.0, 0.5, 0.0]
    eng = lq.Engine(lat, beta, weights=tuple(v), site_weight=gamma / 2, seed=5)
    eng.sweep_many(400, collect=False)
    out = eng.sweep_many(640)
    e = out["ene"] / L
    exact = _tfi_chain_energy(1.0, gamma, 1 / beta)
    err = _berr(e, nb=16)
    assert abs(e.mean() - exact) < 5 * err + 2e-5, (e.mean(), exact, err)   # 2e-5: 1/L^2-size corrections
    spins, ops = eng.get_state()
    assert len(ops) > 900000 and ((ops["loc"] & 1) == 0).sum() > 300000
    ref_labels, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    labels, nc, coll = eng.build_clusters()
    assert nc == ref_nc and np.array_equal(labels, ref_labels)
    for f in SUMS:
        assert coll[f] == pytest.approx(ref[f], rel=1e-7, abs=1e-6), f
    eng.close()


@pytest.mark.parametrize("row", [0, 1, 2])
def test_ladder_observables_vs_exact_diagonalisation(row):
    """4 x 2 ladder (the smallest two-dimensional lattice; tests/golden/ed_ladder.json): Heisenberg,
    XXZ + transverse field, transverse-field Ising against exact diagonalisation."""
    import looper_b200 as lq
    ed = json.load(open(os.path.join(HERE, "golden", "ed_ladder.json")))[row]
    lat = lq.hypercubic_lattice((4, 2))
    assert sorted(zip(lat["src"].tolist(), lat["dst"].tolist())) == sorted(map(tuple, ed["bonds"]))
    n, beta = ed["n"], 1 / ed["T"]
    v, off, sign = lq.xxz_weights(ed["jxy"], ed["jz"])
    eng = lq.Engine(lat, beta, weights=tuple(v), site_weight=ed["gamma"] / 2, seed=777 + row)
    eng.sweep_many(3000, collect=False)
    out = eng.sweep_many(24000)
    eng.close()
    series = {
        "energy_density": out["ene"] / n,
        "umag2": out["umag2"],
        "smag2": out["smag2"],
        "usus_density": beta * out["umag"] / n,
        "ssus_density": beta * out["smag"] / n,
    }
    if v[2] == 0 and v[3] == 0 and ed["gamma"] > 0:
        series["transmag_density"] = 0.5 * out["tlen"] / n
    for k, x in series.items():
        err = _berr(x)
        assert abs(x.mean() - ed[k]) < 4.5 * err + 1e-10, (k, x.mean(), ed[k], err)
