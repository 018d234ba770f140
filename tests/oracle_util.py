"""ctypes access to oracle/liboracle.so -- the CPU checker (tests only; see oracle/oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
OP_DTYPE = np.dtype([("time", "<f8"), ("loc", "<i4"), ("type", "<i4")])
COLL_FIELDS = ["nop", "nc", "noc", "ene",
               "umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
               "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag",
               "sa_usus", "sa_smag", "sa_ssus", "tlen"]


class OrcCollector(C.Structure):
    _fields_ = [(f, C.c_double) for f in COLL_FIELDS]

    def as_dict(self):
        return {f: getattr(self, f) for f in COLL_FIELDS}


def build():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "oracle.cpp")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so", "oracle_loop"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                 C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_looper_estimators.argtypes = [C.c_void_p, C.c_int]
        L.orc_poisson_replay.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_sweep.argtypes = [C.c_void_p, C.POINTER(OrcCollector)]
        L.orc_num_ops.argtypes = [C.c_void_p]
        L.orc_num_ops.restype = C.c_int64
        L.orc_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_get_last_graph.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_build_clusters.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.POINTER(C.c_int64), C.POINTER(OrcCollector)]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_double, C.c_uint32]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_sweep.argtypes = [C.c_void_p, C.POINTER(OrcCollector)]
        L.orc_sse_sweep.argtypes = [C.c_void_p, C.POINTER(OrcCollector)]
        L.orc_sse_collect.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.POINTER(OrcCollector)]
        L.orc_model_num_ops.argtypes = [C.c_void_p]
        L.orc_model_num_ops.restype = C.c_int64
        L.orc_model_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stiffness.restype = C.c_double
        L.orc_stiffness.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
        L.orc_union_find_replay.argtypes = [C.c_char_p, C.c_int]
        L.orc_xxz_weights.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_run_chain.argtypes = [C.c_int, C.c_double, C.c_uint, C.c_uint, C.POINTER(C.c_double)]
        _lib = L
    return _lib


class OracleSim:
    """standalone/loop.C on an arbitrary bond table."""

    def __init__(self, lattice, beta, seed=29833, looper_estimators=True):
        self.N = int(lattice["num_sites"])
        self.src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
        self.dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
        self.B = len(self.src)
        g = lattice.get("gauge")
        self.gauge = np.ascontiguousarray(g if g is not None else np.zeros(self.N), dtype=np.float64)
        self.beta = beta
        self.h = lib().orc_create(self.N, self.B, self.src.ctypes.data, self.dst.ctypes.data,
                                  self.gauge.ctypes.data, beta, seed)
        if not looper_estimators:     # standalone/loop.C statements only (bench.py CPU legs)
            lib().orc_set_looper_estimators(self.h, 0)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def sweep(self):
        c = OrcCollector()
        lib().orc_sweep(self.h, C.byref(c))
        return c.as_dict()

    def get_state(self):
        n = lib().orc_num_ops(self.h)
        spins = np.zeros(self.N, dtype=np.int32)
        ops = np.zeros(n, dtype=OP_DTYPE)
        lib().orc_get_state(self.h, spins.ctypes.data, ops.ctypes.data)
        return spins, ops

    def set_state(self, spins, ops):
        spins = np.ascontiguousarray(spins, dtype=np.int32)
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        lib().orc_set_state(self.h, spins.ctypes.data, ops.ctypes.data, len(ops))

    def last_graph(self):
        n = lib().orc_num_ops(self.h)
        sb = np.zeros(self.N, dtype=np.int32)
        ops = np.zeros(n, dtype=OP_DTYPE)
        lo = np.zeros(n, dtype=np.int32)
        up = np.zeros(n, dtype=np.int32)
        sid = np.zeros(self.N, dtype=np.int32)
        nc = C.c_int32(0)
        flip = np.zeros(self.N + n + 1, dtype=np.int32)
        lib().orc_get_last_graph(self.h, sb.ctypes.data, ops.ctypes.data, lo.ctypes.data,
                                 up.ctypes.data, sid.ctypes.data, C.byref(nc), flip.ctypes.data)
        return dict(spins_before=sb, ops=ops, lower=lo, upper=up, site=sid, nc=nc.value,
                    flip=flip[:nc.value])


def poisson_replay(mean=3.0, count=1 << 20, nbins=15):
    """orc_poisson_replay: (text of test/poisson_distribution.op, histogram)."""
    n = lib().orc_poisson_replay(mean, count, None, 0, None, 0)
    buf = C.create_string_buffer(n + 1)
    bins = np.zeros(nbins, dtype=np.int64)
    lib().orc_poisson_replay(mean, count, buf, n + 1, bins.ctypes.data, nbins)
    return buf.value.decode(), bins


def xxz_weights(jxy, jz, a=0.0):
    """orc_xxz_weights (weight_impl.h:165-188): (v[4], offset, sign)."""
    v = (C.c_double * 4)()
    off, sg = C.c_double(0), C.c_int(0)
    lib().orc_xxz_weights(jxy, jz, a, v, C.byref(off), C.byref(sg))
    return list(v), off.value, sg.value


class OracleModelSim:
    """path_integral.C on an arbitrary bond table with XXZ bond graphs and site graphs."""

    def __init__(self, lattice, beta, weights=(0.5, 0, 0, 0), site_weight=0.0, seed=29833):
        self.N = int(lattice["num_sites"])
        self.src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
        self.dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
        self.B = len(self.src)
        g = lattice.get("gauge")
        self.gauge = np.ascontiguousarray(g if g is not None else np.zeros(self.N), dtype=np.float64)
        self.bw = np.ascontiguousarray(np.tile(np.asarray(weights, dtype=np.float64), self.B))
        self.sw = np.full(self.N, float(site_weight))
        self.beta = beta
        self.h = lib().orc_model_create(self.N, self.B, self.src.ctypes.data, self.dst.ctypes.data,
                                        self.gauge.ctypes.data, self.bw.ctypes.data,
                                        self.sw.ctypes.data, beta, seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_destroy(self.h)
            self.h = None

    def sweep(self):
        c = OrcCollector()
        rc = lib().orc_model_sweep(self.h, C.byref(c))
        if rc != 0:
            raise ValueError(f"orc_model_sweep failed with {rc}")
        return c

    def sse_sweep(self):
        """one step of the SSE worker (sse.C:168-407) on this model"""
        c = OrcCollector()
        rc = lib().orc_sse_sweep(self.h, C.byref(c))
        if rc != 0:
            raise ValueError(f"orc_sse_sweep failed with {rc}")
        return c

    def get_state(self):
        n = lib().orc_model_num_ops(self.h)
        spins = np.zeros(self.N, dtype=np.int32)
        ops = np.zeros(n, dtype=OP_DTYPE)
        lib().orc_model_get_state(self.h, spins.ctypes.data, ops.ctypes.data)
        return spins, ops


def sse_collect(lattice, spins, ops):
    """orc_sse_collect: SSE collector (string positions as times) of a configuration."""
    N = int(lattice["num_sites"])
    src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
    g = lattice.get("gauge")
    gauge = np.ascontiguousarray(g if g is not None else np.zeros(N), dtype=np.float64)
    spins = np.ascontiguousarray(spins, dtype=np.int32)
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    c = OrcCollector()
    rc = lib().orc_sse_collect(N, len(src), src.ctypes.data, dst.ctypes.data, gauge.ctypes.data,
                               spins.ctypes.data, ops.ctypes.data, len(ops), C.byref(c))
    if rc != 0:
        raise ValueError(f"orc_sse_collect failed with {rc}")
    return c.as_dict()


def stiffness(lattice, spins, ops):
    """orc_stiffness: (w2 improved, w2 normal) of a configuration."""
    src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
    vec = np.ascontiguousarray(lattice["bond_vectors"], dtype=np.float64).reshape(-1)
    spins = np.ascontiguousarray(spins, dtype=np.int32)
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    wn = C.c_double(0)
    w2 = lib().orc_stiffness(int(lattice["num_sites"]), len(src), src.ctypes.data, dst.ctypes.data,
                             vec.ctypes.data, int(lattice["vector_dim"]), spins.ctypes.data,
                             ops.ctypes.data, len(ops), C.byref(wn))
    if w2 < 0:
        raise ValueError("orc_stiffness: illegal configuration")
    return w2, wn.value


def build_clusters(lattice, spins, ops):
    """orc_build_clusters: canonical labels (N + 2n), nc, collector dict; raises on bad input."""
    N = int(lattice["num_sites"])
    src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
    g = lattice.get("gauge")
    gauge = np.ascontiguousarray(g if g is not None else np.zeros(N), dtype=np.float64)
    spins = np.ascontiguousarray(spins, dtype=np.int32)
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    n = len(ops)
    labels = np.zeros(N + 2 * n, dtype=np.int32)
    nc = C.c_int64(0)
    c = OrcCollector()
    rc = lib().orc_build_clusters(N, len(src), src.ctypes.data, dst.ctypes.data, gauge.ctypes.data,
                                  spins.ctypes.data, ops.ctypes.data, n, labels.ctypes.data,
                                  C.byref(nc), C.byref(c))
    if rc != 0:
        raise ValueError(f"orc_build_clusters failed with {rc}")
    return labels, nc.value, c.as_dict()
