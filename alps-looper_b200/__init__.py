"""alps-looper_b200 -- Python (ctypes) binding of the C ABI in include/lq.h.

This is plumbing for tests and bench.py: every call goes straight into liblq.so (hand-written
sm_100a CUDA behind `extern "C"`).  There is NO CPU fallback: importing works without a GPU (so
that the symbol table can be checked), but `Engine(...)` raises if the library cannot find a
CUDA device, and the import itself raises if liblq.so has not been built.

The host-side C++ mirror of the reference's worker API lives in alps-looper_b200/looper/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQ_LIB") or os.path.join(_HERE, "liblq.so")   # (LQ_LIB: A/B builds of the kernels, experiments only)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C alps-looper_b200/csrc` "
        "(or __graft_entry__.build()); there is no fallback implementation")

lib = C.CDLL(LIB_PATH)


class LqLattice(C.Structure):
    _fields_ = [("num_sites", C.c_int32), ("num_bonds", C.c_int32),
                ("src", C.POINTER(C.c_int32)), ("dst", C.POINTER(C.c_int32)),
                ("gauge", C.POINTER(C.c_double)), ("dims", C.c_int32 * 3),
                ("vector_dim", C.c_int32), ("bond_vectors", C.POINTER(C.c_double))]


class LqModel(C.Structure):
    _fields_ = [("bond_weights", C.POINTER(C.c_double)), ("uniform_weights", C.c_double * 4),
                ("energy_offset", C.c_double), ("site_weights", C.POINTER(C.c_double)),
                ("uniform_site_weight", C.c_double)]


class LqOptions(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("device", C.c_int32), ("tile_sites", C.c_int32),
                ("window_ops", C.c_double), ("reserve", C.c_double),
                ("cluster_reserve", C.c_double), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("flags", C.c_int32), ("representation", C.c_int32), ("cut", C.c_int32)]


class LqOp(C.Structure):
    _fields_ = [("time", C.c_double), ("loc", C.c_int32), ("type", C.c_int32)]


OP_DTYPE = np.dtype([("time", "<f8"), ("loc", "<i4"), ("type", "<i4")])

COLLECTOR_FIELDS = ["nop", "nc", "noc", "ene",
                    "umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
                    "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag", "tlen", "w2"]


class LqCollector(C.Structure):
    _fields_ = [(f, C.c_double) for f in COLLECTOR_FIELDS]

    def as_dict(self):
        return {f: getattr(self, f) for f in COLLECTOR_FIELDS}


COLLECTOR_DTYPE = np.dtype([(f, "<f8") for f in COLLECTOR_FIELDS])


class LqTiling(C.Structure):
    _fields_ = [("num_tiles", C.c_int32), ("num_classes", C.c_int32), ("max_bonds", C.c_int32),
                ("max_sites", C.c_int32), ("max_halo_buckets", C.c_int32), ("max_walk_halo", C.c_int32),
                ("max_ksites", C.c_int32), ("max_degree", C.c_int32), ("owned_bonds", C.c_int64),
                ("halo_buckets", C.c_int64)]


class LqSpacePlan(C.Structure):
    _fields_ = [("owned_tiles", C.c_int32), ("walked_ghost_tiles", C.c_int32), ("ghost_tiles", C.c_int32),
                ("owned_sites", C.c_int32), ("walked_sites", C.c_int32), ("local_sites", C.c_int32),
                ("segments", C.c_int32), ("neighbours", C.c_int32),
                ("owner_bonds", C.c_int64), ("owner_sites", C.c_int64), ("user_bonds", C.c_int64),
                ("user_sites", C.c_int64), ("checksum_owner", C.c_int64), ("checksum_user", C.c_int64)]


class LqTimer(C.Structure):
    _fields_ = [("id", C.c_int32), ("count", C.c_int32), ("seconds", C.c_double),
                ("label", C.c_char * 40)]


class LqInfo(C.Structure):
    _fields_ = [("num_tiles", C.c_int32), ("num_windows", C.c_int32),
                ("page_capacity", C.c_int32), ("threads_per_page", C.c_int32),
                ("op_capacity", C.c_int64), ("cluster_capacity", C.c_int64),
                ("device_bytes", C.c_int64), ("sm_count", C.c_int32), ("nodes_per_op", C.c_int32)]


ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
ALL_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
SEND_RECV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                           C.c_void_p)


class LqComm(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("all_gather", ALL_GATHER_FN),
                ("all_reduce_i64", ALL_REDUCE_FN), ("send_recv", SEND_RECV_FN)]


# every symbol include/lq.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = ["lq_create", "lq_destroy", "lq_set_beta", "lq_set_state", "lq_get_state", "lq_get_step", "lq_set_step", "lq_sweep",
           "lq_sweep_many", "lq_build_clusters", "lq_timers", "lq_enable_timers", "lq_get_info", "lq_tiling_info", "lq_space_plan_info", "lq_kernel_launches", "lq_regrow_count", "lq_h2d_bytes", "lq_d2h_bytes",
           "lq_set_comm", "lq_comm_unique_id", "lq_comm_init", "lq_stream", "lq_last_error", "lq_version"]

_h = C.c_void_p
lib.lq_create.argtypes = [C.POINTER(_h), C.POINTER(LqLattice), C.POINTER(LqModel), C.c_double,
                          C.POINTER(LqOptions)]
lib.lq_destroy.argtypes = [_h]
lib.lq_set_beta.argtypes = [_h, C.c_double]
lib.lq_set_state.argtypes = [_h, C.c_void_p, C.c_void_p, C.c_int64]
lib.lq_get_state.argtypes = [_h, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
lib.lq_sweep.argtypes = [_h, C.POINTER(LqCollector)]
lib.lq_sweep_many.argtypes = [_h, C.c_int32, C.c_void_p]
lib.lq_build_clusters.argtypes = [_h, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(LqCollector)]
lib.lq_timers.argtypes = [_h, C.POINTER(LqTimer), C.POINTER(C.c_int32)]
lib.lq_enable_timers.argtypes = [_h, C.c_int]
lib.lq_get_info.argtypes = [_h, C.POINTER(LqInfo)]
lib.lq_kernel_launches.argtypes = [_h]
lib.lq_kernel_launches.restype = C.c_int64
lib.lq_get_step.argtypes = [_h]
lib.lq_get_step.restype = C.c_uint32
lib.lq_set_step.argtypes = [_h, C.c_uint32]
lib.lq_regrow_count.argtypes = [_h]
lib.lq_regrow_count.restype = C.c_int64
lib.lq_h2d_bytes.argtypes = [_h]
lib.lq_h2d_bytes.restype = C.c_int64
lib.lq_d2h_bytes.argtypes = [_h]
lib.lq_d2h_bytes.restype = C.c_int64
lib.lq_set_comm.argtypes = [_h, C.POINTER(LqComm)]
lib.lq_comm_unique_id.argtypes = [C.c_void_p]
lib.lq_comm_init.argtypes = [_h, C.c_void_p, C.c_int32, C.c_int32]
lib.lq_stream.argtypes = [_h]
lib.lq_stream.restype = C.c_void_p
lib.lq_last_error.restype = C.c_char_p
lib.lq_version.restype = C.c_char_p


class LqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lq error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise LqError(rc, lib.lq_last_error().decode())


# --------------------------------------------------------------------------------------------
# lattice helpers (host logic; the hypercubic generator mirrors what looper::lattice_helper
# hands to the worker: bonds with source/target and the bipartite gauge, lattice.h:85,290-371)
# --------------------------------------------------------------------------------------------
def chain_lattice(L):
    """standalone/common.h:92-93: bond b joins b -- (b+1) mod L."""
    src = np.arange(L, dtype=np.int32)
    dst = ((src + 1) % L).astype(np.int32)
    gauge = np.where(np.arange(L) % 2 == 0, 1.0, -1.0)
    vec = np.zeros((L, 3))
    vec[:, 0] = 1.0 / L     # ALPS bond_vector_relative: bond vector over the lattice extent
    return dict(num_sites=L, src=src, dst=dst, gauge=gauge, dims=(L, 0, 0), bond_vectors=vec, vector_dim=1)


def hypercubic_lattice(dims):
    """Periodic hypercubic lattice; site = x + L0*(y + L1*z); bonds of direction k first for all
    sites, then direction k+1 (SURVEY Appendix C '2-D baseline')."""
    dims = [int(d) for d in dims if int(d) > 1]
    n = int(np.prod(dims))
    idx = np.arange(n)
    coords = []
    rem = idx.copy()
    for d in dims:
        coords.append(rem % d)
        rem = rem // d
    src, dst, vecs = [], [], []
    stride = 1
    for k, d in enumerate(dims):
        nxt = idx + stride * (((coords[k] + 1) % d) - coords[k])
        if d == 2:
            # a periodic ring of length 2 has ONE bond per pair in this direction
            keep = coords[k] == 0
            src.append(idx[keep]); dst.append(nxt[keep])
        else:
            src.append(idx); dst.append(nxt)
        v = np.zeros((len(src[-1]), 3))   # relative bond vector: 1/extent along direction k (also across the seam)
        if k < 3:
            v[:, k] = 1.0 / d
        vecs.append(v)
        stride *= d
    src = np.concatenate(src).astype(np.int32)
    dst = np.concatenate(dst).astype(np.int32)
    parity = np.zeros(n, dtype=np.int64)
    for c in coords:
        parity += c
    bip = all(d % 2 == 0 for d in dims)
    gauge = np.where(parity % 2 == 0, 1.0, -1.0) if bip else np.zeros(n)
    dd = tuple(dims + [0] * (3 - len(dims)))
    return dict(num_sites=n, src=src, dst=dst, gauge=gauge, dims=dd,
                bond_vectors=np.concatenate(vecs), vector_dim=min(len(dims), 3))


def tiling_info(lattice, tile_sites=0, with_sites=False):
    """lq_tiling_info: the host-side spatial tiling of a lattice (runs without a GPU)."""
    src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
    lat = LqLattice()
    lat.num_sites = int(lattice["num_sites"])
    lat.num_bonds = len(src)
    lat.src = src.ctypes.data_as(C.POINTER(C.c_int32))
    lat.dst = dst.ctypes.data_as(C.POINTER(C.c_int32))
    lat.gauge = None
    lat.dims = (C.c_int32 * 3)(*tuple(lattice.get("dims", (0, 0, 0))))
    lat.vector_dim = 0
    lat.bond_vectors = None
    out = LqTiling()
    _check(lib.lq_tiling_info(C.byref(lat), int(tile_sites), 1 if with_sites else 0, C.byref(out)))
    return {f: getattr(out, f) for f, _ in LqTiling._fields_}


def space_plan_info(lattice, nranks, rank, tile_sites=0, with_sites=False, peer=-1):
    """lq_space_plan_info: what rank `rank` of `nranks` owns and mirrors under the spatial cut (no GPU)."""
    src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
    lat = LqLattice()
    lat.num_sites = int(lattice["num_sites"])
    lat.num_bonds = len(src)
    lat.src = src.ctypes.data_as(C.POINTER(C.c_int32))
    lat.dst = dst.ctypes.data_as(C.POINTER(C.c_int32))
    lat.gauge = None
    lat.dims = (C.c_int32 * 3)(*tuple(lattice.get("dims", (0, 0, 0))))
    lat.vector_dim = 0
    lat.bond_vectors = None
    out = LqSpacePlan()
    lib.lq_space_plan_info.argtypes = [C.POINTER(LqLattice), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(LqSpacePlan)]
    _check(lib.lq_space_plan_info(C.byref(lat), int(tile_sites), 1 if with_sites else 0, int(nranks), int(rank),
                                  int(peer), C.byref(out)))
    return {f: getattr(out, f) for f, _ in LqSpacePlan._fields_}


def xxz_weights(jxy, jz, a=0.0):
    """looper/weight_impl.h:165-188 (standard / 'ergodic' solution); returns (v[4], offset, sign)."""
    sign = 1 if jxy <= 0 else -1
    jxy = abs(jxy)
    a = min(max(a, 0.0), 1.0)
    crop0 = lambda x: x if x > 0 else 0.0
    if jxy + abs(jz) > 1e-10:
        if jxy - jz > 2 * a * jxy:
            v = [crop0(min(jxy / 2, (jxy + jz) / 4)), crop0(min(jxy / 2, (jxy - jz) / 4)),
                 crop0(-(jxy - jz) / 2.0), crop0(-(jxy + jz) / 2)]
        else:
            v = [(1 - a) * jxy / 2, a * jxy / 2, -((1 - 2 * a) * jxy - jz) / 2, 0.0]
    else:
        v = [0.0] * 4
    return v, sum(v) / 2, sign


class Engine:
    """One loop-update engine on one GPU (owns spins + operator string, like loop_worker)."""

    def __init__(self, lattice, beta, weights=(0.5, 0.0, 0.0, 0.0), energy_offset=None, seed=29833,
                 device=0, tile_sites=0, window_ops=0.0, reserve=0.0, cluster_reserve=0.0,
                 rank=0, nranks=1, timers=False, bond_weights=None, site_weight=0.0, site_weights=None,
                 stiffness=False, sse=False, cut="time"):
        self.lattice = lattice
        self.N = int(lattice["num_sites"])
        self._src = np.ascontiguousarray(lattice["src"], dtype=np.int32)
        self._dst = np.ascontiguousarray(lattice["dst"], dtype=np.int32)
        self.B = len(self._src)
        g = lattice.get("gauge")
        self._gauge = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        lat = LqLattice()
        lat.num_sites = self.N
        lat.num_bonds = self.B
        lat.src = self._src.ctypes.data_as(C.POINTER(C.c_int32))
        lat.dst = self._dst.ctypes.data_as(C.POINTER(C.c_int32))
        lat.gauge = (self._gauge.ctypes.data_as(C.POINTER(C.c_double))
                     if self._gauge is not None else None)
        dims = tuple(lattice.get("dims", (0, 0, 0)))
        lat.dims = (C.c_int32 * 3)(*dims)
        # winding-number estimator (stiffness.h): only when asked for and the lattice has vectors
        self._bvec = None
        lat.vector_dim = 0
        lat.bond_vectors = None
        if stiffness:
            if lattice.get("bond_vectors") is None:
                raise ValueError("stiffness needs lattice['bond_vectors'] (B x 3) and lattice['vector_dim']")
            self._bvec = np.ascontiguousarray(lattice["bond_vectors"], dtype=np.float64).reshape(-1)
            assert self._bvec.size == 3 * self.B
            lat.vector_dim = int(lattice.get("vector_dim", 3))
            lat.bond_vectors = self._bvec.ctypes.data_as(C.POINTER(C.c_double))
        self.vector_dim = lat.vector_dim
        mod = LqModel()
        self._bw = None
        if bond_weights is not None:
            self._bw = np.ascontiguousarray(bond_weights, dtype=np.float64).reshape(-1)
            assert self._bw.size == 4 * self.B
            mod.bond_weights = self._bw.ctypes.data_as(C.POINTER(C.c_double))
            wsum = float(self._bw.sum())
        else:
            mod.bond_weights = None
            wsum = float(sum(weights)) * self.B
        # site graph weights v0 = |Hx|/2 (weight_impl.h:62-88); their offsets (= v0 each) join the
        # energy offset like the bond offsets do
        self._sw = None
        if site_weights is not None:
            self._sw = np.ascontiguousarray(site_weights, dtype=np.float64).reshape(-1)
            assert self._sw.size == self.N
            mod.site_weights = self._sw.ctypes.data_as(C.POINTER(C.c_double))
            swsum = float(self._sw.sum())
        else:
            mod.site_weights = None
            swsum = float(site_weight) * self.N
        mod.uniform_site_weight = float(site_weight)
        mod.uniform_weights = (C.c_double * 4)(*weights)
        # energy_offset = sum of bond offsets = sum of weights / 2 (weight_impl.h:187)
        mod.energy_offset = wsum / 2 + swsum if energy_offset is None else energy_offset
        self.energy_offset = mod.energy_offset
        opt = LqOptions(seed=seed, device=device, tile_sites=tile_sites, window_ops=window_ops,
                        reserve=reserve, cluster_reserve=cluster_reserve, rank=rank,
                        nranks=nranks, flags=1 if timers else 0, representation=1 if sse else 0,
                        cut={"time": 0, "space": 1}[cut])
        self.sse = bool(sse)
        self.beta = float(beta)
        self._h = _h()
        _check(lib.lq_create(C.byref(self._h), C.byref(lat), C.byref(mod), self.beta, C.byref(opt)))
        self._comm_keep = None

    def close(self):
        if getattr(self, "_h", None):
            lib.lq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = LqInfo()
        _check(lib.lq_get_info(self._h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in LqInfo._fields_}

    def sweep(self):
        c = LqCollector()
        _check(lib.lq_sweep(self._h, C.byref(c)))
        return c.as_dict()

    def sweep_many(self, count, collect=True):
        if collect:
            out = np.zeros(count, dtype=COLLECTOR_DTYPE)
            _check(lib.lq_sweep_many(self._h, count, out.ctypes.data))
            return out
        _check(lib.lq_sweep_many(self._h, count, None))
        return None

    def set_beta(self, beta):
        _check(lib.lq_set_beta(self._h, float(beta)))
        self.beta = float(beta)

    def num_ops(self):
        n = C.c_int64(0)
        _check(lib.lq_get_state(self._h, None, None, C.byref(n)))
        return n.value

    def get_state(self):
        n = self.num_ops()
        spins = np.zeros(self.N, dtype=np.int32)
        ops = np.zeros(n, dtype=OP_DTYPE)
        nn = C.c_int64(n)
        _check(lib.lq_get_state(self._h, spins.ctypes.data, ops.ctypes.data if n else None,
                                C.byref(nn)))
        return spins, ops

    def set_state(self, spins, ops):
        spins = np.ascontiguousarray(spins, dtype=np.int32)
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        _check(lib.lq_set_state(self._h, spins.ctypes.data, ops.ctypes.data if len(ops) else None,
                                len(ops)))

    def build_clusters(self):
        n = self.num_ops()
        labels = np.zeros(self.N + 2 * n, dtype=np.int32)
        nc = C.c_int64(0)
        c = LqCollector()
        _check(lib.lq_build_clusters(self._h, labels.ctypes.data, C.byref(nc), C.byref(c)))
        return labels, nc.value, c.as_dict()

    def timers(self):
        cnt = C.c_int32(32)
        arr = (LqTimer * 32)()
        _check(lib.lq_timers(self._h, arr, C.byref(cnt)))
        return [dict(id=arr[i].id, count=arr[i].count, seconds=arr[i].seconds,
                     label=arr[i].label.decode()) for i in range(min(cnt.value, 32))]

    def enable_timers(self, on=True):
        _check(lib.lq_enable_timers(self._h, 1 if on else 0))

    def kernel_launches(self):
        return int(lib.lq_kernel_launches(self._h))

    def get_step(self):
        return int(lib.lq_get_step(self._h))

    def set_step(self, step):
        _check(lib.lq_set_step(self._h, int(step)))

    def regrow_count(self):
        return int(lib.lq_regrow_count(self._h))

    def copied_bytes(self):
        return int(lib.lq_h2d_bytes(self._h)), int(lib.lq_d2h_bytes(self._h))

    def stream(self):
        return lib.lq_stream(self._h)

    def comm_init(self, unique_id, rank, nranks):
        """lq_comm_init: the engine's own NCCL communicator (unique_id: 128 bytes from
        `nccl_unique_id()` on one rank, distributed by the host)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(lib.lq_comm_init(self._h, buf, int(rank), int(nranks)))

    def set_comm(self, all_gather, all_reduce_i64, send_recv=None):
        ag = ALL_GATHER_FN(all_gather)
        ar = ALL_REDUCE_FN(all_reduce_i64)
        sr = SEND_RECV_FN(send_recv) if send_recv is not None else SEND_RECV_FN()   # (NULL: imaginary-time slabs only)
        comm = LqComm(ctx=None, all_gather=ag, all_reduce_i64=ar, send_recv=sr)
        self._comm_keep = (ag, ar, sr, comm)
        _check(lib.lq_set_comm(self._h, C.byref(comm)))


def nccl_unique_id():
    """lq_comm_unique_id: 128 bytes for lq_comm_init, to be created on ONE rank."""
    buf = (C.c_char * 128)()
    _check(lib.lq_comm_unique_id(buf))
    return bytes(buf)


def observables(coll, beta, num_sites, sse=False):
    """collector.commit() of energy.h:74-83 and susceptibility.h:199-254 (path-integral branch):
    the named scalar observables of one Monte Carlo step."""
    vol = float(num_sites)
    o = {}
    o["Energy"] = coll["ene"]
    o["Energy Density"] = coll["ene"] / vol
    o["Energy^2"] = coll["ene"] ** 2 - coll["nop"] / beta ** 2
    o["Number of Clusters"] = coll["nc"]
    o["|Magnetization|"] = abs(coll["umag0"])
    o["Magnetization^2"] = coll["umag2"]
    o["Magnetization^4"] = 3 * coll["umag2"] ** 2 - 2 * coll["umag4"]
    nop = coll["nop"]
    dip = (lambda x: x / nop if nop > 0 else 0.0)          # divide_if_positive.h
    # susceptibility.h:213-215: SSE strings carry integer times 0..n-1 and the top n
    sus = (lambda x, x2: beta * (dip(x) + x2) / (nop + 1) / vol) if sse else (lambda x, x2: beta * x / vol)
    o["Susceptibility"] = sus(coll["umag"], coll["umag2"])
    o["Generalized Magnetization^2"] = coll["usize2"]
    o["Generalized Susceptibility"] = sus(coll["usize"], coll["usize2"])
    o["|Staggered Magnetization|"] = abs(coll["smag0"])
    o["Staggered Magnetization^2"] = coll["smag2"]
    o["Staggered Magnetization^4"] = 3 * coll["smag2"] ** 2 - 2 * coll["smag4"]
    o["Staggered Susceptibility"] = sus(coll["smag"], coll["smag2"])
    o["Generalized Staggered Magnetization^2"] = coll["ssize2"]
    o["Generalized Staggered Susceptibility"] = sus(coll["ssize"], coll["ssize2"])
    # transmag.h:106-109
    o["Transverse Magnetization"] = 0.5 * coll["tlen"]
    o["Transverse Magnetization Density"] = 0.5 * coll["tlen"] / vol
    return o


def stiffness(coll, beta, vector_dim):
    """stiffness.h:131-133: "Stiffness" = w2 / (beta * dim)."""
    return coll["w2"] / (beta * vector_dim)
