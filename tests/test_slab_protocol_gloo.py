"""CPU, world_size 2, gloo: the multi-rank exchange protocol of the slab engine
(csrc/lq_kernels.cuh k_mr_*; reference: looper/parallel.h:1609-1809) restated with numpy on the
host and run across two real processes: every rank labels its imaginary-time slab, publishes the
2N boundary ids, all-gathers them, unifies top(r) with bottom(r+1) redundantly and all-reduces the
open-cluster partial sums.  The merged cluster count and the sum of squared cluster lengths must
equal the oracle's on the whole configuration."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_util as orc          # noqa: E402
import looper_lattices as ll       # noqa: E402


def find(p, x):
    while p[x] != x:
        p[x] = p[p[x]]
        x = p[x]
    return x


def union(p, a, b):
    ra, rb = find(p, a), find(p, b)
    if ra != rb:
        if ra < rb:
            ra, rb = rb, ra
        p[ra] = rb                      # larger root under smaller: min-index roots


def slab_label(lat, spins0, ops, t0, t1):
    """local labelling of the slab [t0,t1): nodes 0..N-1 bottom boundary, N+k operators."""
    N = lat["num_sites"]
    sel = ops[(ops["time"] >= t0) & (ops["time"] < t1)]
    parent = list(range(N + len(sel)))
    cur = list(range(N))
    length = np.zeros(N + len(sel))
    for k, o in enumerate(sel):
        b = o["loc"] >> 1
        s0, s1 = int(lat["src"][b]), int(lat["dst"][b])
        union(parent, cur[s0], cur[s1])
        length[cur[s0]] += 2 * o["time"]     # standalone/loop.C:143-146, per node instead of per id
        cur[s0] = cur[s1] = N + k
        length[N + k] -= 2 * o["time"]
    for s in range(N):                       # slab boundaries (loop_mpi.C:155-182)
        length[s] -= t0
        length[cur[s]] += t1
    roots = [find(parent, x) for x in range(len(parent))]
    csum = {}
    for x, r in enumerate(roots):
        csum[r] = csum.get(r, 0.0) + length[x]
    return roots, cur, csum, len(sel)


def worker(rank, world, port, lat, spins, ops, ref_nc, ref_ssus, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = lat["num_sites"]
    roots, cur, csum, nloc = slab_label(lat, spins, ops, rank / world, (rank + 1) / world)
    # open ids: bottom-touching clusters by their root site (< N), top-only ones by N + min top site
    topmin = {}
    for s in range(N):
        r = roots[cur[s]]
        if r >= N:
            topmin[r] = min(topmin.get(r, s), s)
    oid = lambda r: r if r < N else N + topmin[r]
    send = torch.tensor([oid(roots[s]) for s in range(N)] + [oid(roots[cur[s]]) for s in range(N)], dtype=torch.int64)
    recv = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    gp = list(range(world * 2 * N))
    used = set()
    for r in range(world):
        rn = (r + 1) % world
        for s in range(N):
            a, b = r * 2 * N + int(recv[r][N + s]), rn * 2 * N + int(recv[rn][s])
            used.update((a, b))
            union(gp, a, b)
    groots = sorted({find(gp, x) for x in used})
    gid = {g: i for i, g in enumerate(groots)}
    # partial sums of my open clusters -> global table -> all-reduce
    table = torch.zeros(len(groots), dtype=torch.float64)
    open_roots = {roots[s] for s in range(N)} | {roots[cur[s]] for s in range(N)}
    for r in open_roots:
        table[gid[find(gp, rank * 2 * N + oid(r))]] += csum[r]
    dist.all_reduce(table)
    closed = [v for r, v in csum.items() if r not in open_roots]
    loc = torch.tensor([float(len(closed)), float(sum(v * v for v in closed)), float(nloc)], dtype=torch.float64)
    dist.all_reduce(loc)
    nc = int(loc[0]) + len(groots)
    ssus = float(loc[1]) + float((table * table).sum())
    ok = (nc == ref_nc) and abs(ssus - ref_ssus) < 1e-9 * max(1.0, ref_ssus) and int(loc[2]) == len(ops)
    q.put((rank, ok, nc, ssus))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["chain16", "square6"])
def test_two_rank_slab_merge_over_gloo(case):
    lat, beta = (ll.chain_lattice(16), 10.0) if case == "chain16" else (ll.hypercubic_lattice((6, 6)), 4.0)
    sim = orc.OracleSim(lat, beta)
    for _ in range(150):
        sim.sweep()
    spins, ops = sim.get_state()
    _, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=worker, args=(r, 2, port, lat, spins, ops, ref_nc, ref["sa_ssus"], q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, nc, ssus in res:
        assert ok, (rank, nc, ref_nc, ssus, ref["sa_ssus"])
