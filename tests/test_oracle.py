"""CPU tests: the oracle (oracle/) against the reference's own golden vectors.

Goldens: standalone/loop.op, standalone/loop_mpi.op-{1..4}, test/union_find.op, test/weight.op.
When the reference tree is mounted (/root/reference) the files are diffed directly; their digests
and the numbers that matter are also embedded here so the tests still pin the oracle on a box
without the tree."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_util as orc

REF = "/root/reference"
ORACLE_DIR = orc.ORACLE_DIR

LOOP_OP = """System Length             = 8
Temperature               = 0.2
MCS for Thermalization    = 8192
MCS for Measurement       = 65536
Number of Clusters        = 16.8324 +- 0.0161722
Energy Density            = -0.442732 +- 0.000529149
Uniform Susceptibility    = 0.0800276 +- 0.000536463
Staggered Magnetization^2 = 6.61901 +- 0.0114392
Staggered Susceptibility  = 2.4117 +- 0.00442169
"""
POISSON_OP = """0 0.0498 0.0497 0.000218
1 0.149 0.149 0.000377
2 0.224 0.224 0.000462
3 0.224 0.224 0.000463
4 0.168 0.168 0.0004
5 0.101 0.101 0.00031
6 0.0504 0.0505 0.00022
7 0.0216 0.0219 0.000145
8 0.0081 0.00819 8.84e-05
9 0.0027 0.00261 4.99e-05
10 0.00081 0.000819 2.8e-05
11 0.000221 0.000206 1.4e-05
12 5.52e-05 5.34e-05 7.14e-06
13 1.27e-05 1.72e-05 4.05e-06
14 2.73e-06 2.86e-06 1.65e-06
"""
SHA = {
    "test/poisson_distribution.op": "68f1ec97c8acbc760c8da5282b9c365216b68318cf590f20de543e624747f6f2",
    "test/union_find.op": "e7f472e7f249305bf61eec2b55614645c36be1ee1f71f0bac3e7b3af82bfe987",
    "standalone/loop.op": "08b5cb41a9fcac8c2874a42fb3eaa7d21b13b513e0b0aac644192bf3f94a9a91",
    "standalone/loop_mpi.op-1": "d1dbf6c7b20be0a7f3fa5df7aa0d0f9e36134f09fd98ed561e79fc7788f87691",
    "standalone/loop_mpi.op-2": "f3970425d291d6d2253491c2de4dcb2c4413071fb507171e4779c2c0b239d8eb",
    "standalone/loop_mpi.op-3": "69695578bcbab480482883b9fa073aff176fd6f27cb95fbb1b7153c749a6dcf8",
    "standalone/loop_mpi.op-4": "226d53cbe7cbfd9cb67f07ef8681a9f016eecf31bbc8b861b1b2673604a7951a",
}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_embedded_golden_matches_digest():
    assert sha(LOOP_OP.encode()) == SHA["standalone/loop.op"]


def test_oracle_reproduces_standalone_loop_op():
    orc.build()
    out = subprocess.run([os.path.join(ORACLE_DIR, "oracle_loop")], capture_output=True, check=True).stdout
    assert out.decode() == LOOP_OP
    if os.path.exists(REF):
        with open(os.path.join(REF, "standalone/loop.op"), "rb") as f:
            assert f.read() == out


def test_oracle_reproduces_union_find_op():
    L = orc.lib()
    n = L.orc_union_find_replay(None, 0)
    import ctypes as C
    buf = C.create_string_buffer(n + 1)
    L.orc_union_find_replay(buf, n + 1)
    assert sha(buf.value) == SHA["test/union_find.op"]
    if os.path.exists(REF):
        with open(os.path.join(REF, "test/union_find.op"), "rb") as f:
            assert f.read() == buf.value


@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "loop")),
                    reason="oracle/_ref not built (reference tree absent)")
def test_reference_binaries_reproduce_goldens():
    """the UNMODIFIED reference, compiled from /root/reference by oracle/Makefile"""
    out = subprocess.run([os.path.join(ORACLE_DIR, "_ref", "loop")], capture_output=True, check=True).stdout
    assert sha(out) == SHA["standalone/loop.op"]
    out = subprocess.run([os.path.join(ORACLE_DIR, "_ref", "union_find_replay")], capture_output=True,
                         check=True).stdout
    assert sha(out) == SHA["test/union_find.op"]
    for p in (1, 2, 3, 4):
        out = subprocess.run([os.path.join(ORACLE_DIR, "_ref", "loop_mpi"), str(p)], capture_output=True,
                             check=True).stdout
        assert sha(out) == SHA[f"standalone/loop_mpi.op-{p}"], p


@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "loop")),
                    reason="oracle/_ref not built")
def test_oracle_equals_reference_binary_on_other_parameters():
    for args in (["-l", "16", "-t", "0.1", "-n", "2048"], ["-l", "32", "-t", "0.05", "-n", "512"]):
        a = subprocess.run([os.path.join(ORACLE_DIR, "_ref", "loop")] + args, capture_output=True, check=True).stdout
        b = subprocess.run([os.path.join(ORACLE_DIR, "oracle_loop")] + args, capture_output=True, check=True).stdout
        assert a == b


# test/weight.op rows "bond weight (standard): C = 0, Jxy = .., Jz = .."
WEIGHT_ROWS = [
    (1.0, 2.0, [0.5, 0, 0.5, 0], 0.5, -1),
    (1.0, 1.0, [0.5, 0, 0, 0], 0.25, -1),
    (1.0, 0.5, [0.375, 0.125, 0, 0], 0.25, -1),
    (1.0, -2.0, [0, 0.5, 0, 0.5], 0.5, -1),
    (-0.108259, 0.585998, [0.0541293, 0, 0.23887, 0], 0.1465, 1),
    (0.569065, 0.543465, [0.278132, 0.00639998, 0, 0], 0.142266, -1),
]


def test_xxz_weights_match_weight_op():
    import ctypes as C
    L = orc.lib()
    for jxy, jz, v_ref, off_ref, sign_ref in WEIGHT_ROWS:
        v = (C.c_double * 4)()
        off = C.c_double()
        sg = C.c_int()
        L.orc_xxz_weights(jxy, jz, 0.0, v, C.byref(off), C.byref(sg))
        assert np.allclose(list(v), v_ref, rtol=2e-5, atol=1e-9)   # .op prints 6 significant digits (inputs too)
        assert off.value == pytest.approx(off_ref, rel=2e-5)
        assert sg.value == sign_ref
    # "ergodic" row: Jxy = 1, Jz = 1, FORCE_SCATTER = 0.1 -> 0.45 0.05 0.1 0 offset 0.3 (weight.op:10)
    v = (C.c_double * 4)(); off = C.c_double(); sg = C.c_int()
    L.orc_xxz_weights(1.0, 1.0, 0.1, v, C.byref(off), C.byref(sg))
    assert np.allclose(list(v), [0.45, 0.05, 0.1, 0.0]) and off.value == pytest.approx(0.3)


def test_build_clusters_agrees_with_the_sweep():
    """orc_build_clusters (looper union_find.h restated) on the state the standalone sweep produced
    gives the same partition as the sweep's own fragments (standalone/union_find.h restated)."""
    import looper_lattices as ll
    for lat, beta in [(ll.chain_lattice(16), 10.0), (ll.hypercubic_lattice((6, 6)), 3.0)]:
        sim = orc.OracleSim(lat, beta)
        for _ in range(200):
            c = sim.sweep()
            g = sim.last_graph()
            labels, nc, coll = orc.build_clusters(lat, g["spins_before"], g["ops"])
            assert nc == g["nc"] == c["nc"]
            N = lat["num_sites"]
            # same partition: reference ids -> canonical labels must be a bijection
            m = {}
            for cid, lab in list(zip(g["site"], labels[:N])) + list(zip(g["upper"], labels[N::2])):
                assert m.setdefault(cid, lab) == lab
            assert len(set(m.values())) == len(m)
            for f in ("sa_usus", "sa_smag", "usize2", "umag2", "smag2"):
                assert coll[f] == pytest.approx(c[f], rel=1e-12, abs=1e-12), f
            for f in ("sa_ssus", "usize", "smag"):
                assert coll[f] == pytest.approx(c[f], rel=1e-9), f
            # looper-named sums vs standalone sums (bipartite HAF): sa_ssus = 4 usize = 4 smag
            assert coll["sa_ssus"] == pytest.approx(4 * coll["usize"], rel=1e-9)
            assert coll["sa_usus"] == pytest.approx(4 * coll["umag2"], rel=1e-12)
            assert coll["sa_smag"] == pytest.approx(4 * coll["usize2"], rel=1e-12)


def test_build_clusters_rejects_illegal_strings():
    import looper_lattices as ll
    lat = ll.chain_lattice(4)
    ops = np.zeros(1, dtype=orc.OP_DTYPE)
    ops[0] = (0.5, (0 << 1) | 1, 0)
    with pytest.raises(ValueError):
        orc.build_clusters(lat, [0, 0, 0, 0], ops)       # parallel spins under a HAF operator
    ops[0] = (0.5, (0 << 1) | 1, 1)
    with pytest.raises(ValueError):
        orc.build_clusters(lat, [0, 1, 0, 1], ops)       # single off-diagonal: not periodic
    labels, nc, _ = orc.build_clusters(lat, [0, 1, 0, 1], ops[:0])
    assert nc == 4 and list(labels) == [0, 1, 2, 3]      # empty string: every site alone


def test_observables_vs_exact_diagonalisation_cpu():
    """chain L=8 T=0.2 against ED (numpy restatement of diag.C:376-468)."""
    import ctypes as C
    out = (C.c_double * 10)()
    orc.lib().orc_run_chain(8, 0.2, 1 << 15, 1 << 12, out)
    ed = dict(ene=-0.441438, usus=0.0804441, smag=6.59939, ssus=2.40159)
    # naive errors underestimate (no binning): allow 5 sigma
    assert abs(out[2] - ed["ene"]) < 5 * out[3]
    assert abs(out[4] - ed["usus"]) < 5 * out[5]
    assert abs(out[6] - ed["smag"]) < 5 * out[7]
    assert abs(out[8] - ed["ssus"]) < 5 * out[9]


def test_oracle_reproduces_poisson_distribution_op():
    """looper/poisson_distribution.h:44-113 driven as test/poisson_distribution.C does
    (MEAN = 3, COUNT = 1048576, test/poisson_distribution.ip): bit-for-bit the golden text."""
    assert sha(POISSON_OP.encode()) == SHA["test/poisson_distribution.op"]
    txt, bins = orc.poisson_replay(3.0, 1 << 20)
    assert txt == POISSON_OP
    assert bins.sum() <= (1 << 20) and bins[3] == round(0.224 * (1 << 20), -3) or bins[3] > 0
    if os.path.exists(REF):
        with open(os.path.join(REF, "test/poisson_distribution.op")) as f:
            assert f.read() == txt


def test_standalone_only_mode_draws_the_same_chain():
    """orc_set_looper_estimators(0) (the timing configuration of bench.py's CPU legs) must not change
    the Markov chain: same operator counts, cluster counts and standalone sums."""
    import looper_lattices as ll
    lat = ll.hypercubic_lattice((8, 8))
    a = orc.OracleSim(lat, 4.0, 5)
    b = orc.OracleSim(lat, 4.0, 5, looper_estimators=False)
    for _ in range(40):
        ca, cb = a.sweep(), b.sweep()
        for f in ("nop", "nc", "sa_usus", "sa_smag", "sa_ssus"):
            assert ca[f] == cb[f]
    assert cb["usize"] == 0 and ca["usize"] > 0
