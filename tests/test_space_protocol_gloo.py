"""CPU, world_size 2, gloo: the multi-rank protocol of the SPATIAL cut (csrc/lq_space.cuh, lq_engine.cu
SpacePlan; reference analogue: the bond ownership of looper/lattice.h:692-787 + the chunk merge of
looper/parallel.h:1609-1809) restated with numpy on the host and run across two real processes.
Rank r owns the sites [r L / 2, (r+1) L / 2) of a chain and the bonds whose SOURCE it owns; it applies
the edges of its own operators only, on a local forest that also holds ghost copies of the operators
on the foreign bonds touching one of its K-sites (own sites + far ends of owned bonds).  Boundary
entries (the far-end sites and all operators on those foreign bonds, in (bond, slot) order on both
sides) publish the smallest entry of their local cluster; one all-gather; entry i of the owner's
segment is unified with entry i of the user's; open-cluster sums meet in one all-reduce.  The merged
cluster count and the sum of squared cluster lengths must equal the oracle's."""
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
        p[ra] = rb


def local_forest(lat, ops, rank, world):
    """returns (parent, length per node, node of every global op index or -1, boundary entries)"""
    N = lat["num_sites"]
    src, dst = lat["src"], lat["dst"]
    owner_site = lambda s: s * world // N
    owner_bond = lambda b: owner_site(int(src[b]))
    own_sites = [s for s in range(N) if owner_site(s) == rank]
    ksites = set(own_sites) | {int(dst[b]) for b in range(len(src)) if owner_bond(b) == rank}
    # local nodes: K-sites, then every operator that touches a K-site (owned or ghost)
    node_site = {s: i for i, s in enumerate(sorted(ksites))}
    nodes = len(node_site)
    op_node = {}
    parent = list(range(nodes))
    cur = {s: node_site[s] for s in ksites}
    length = [0.0] * nodes
    for k, o in enumerate(ops):
        b = int(o["loc"]) >> 1
        s0, s1 = int(src[b]), int(dst[b])
        if s0 not in ksites and s1 not in ksites:
            continue
        parent.append(len(parent))
        length.append(0.0)
        x = len(parent) - 1
        op_node[k] = x
        if owner_bond(b) == rank:          # the edges and the estimator legs of OWNED operators only
            union(parent, cur[s0], cur[s1])
            length[cur[s0]] += 2 * o["time"]
            length[x] -= 2 * o["time"]
        if s0 in ksites:
            cur[s0] = x
        if s1 in ksites:
            cur[s1] = x
    for s in own_sites:                     # close the world lines of the own sites (k_close) + their ends
        union(parent, node_site[s], cur[s])
        length[cur[s]] += 1.0
    # boundary segments (owner q -> user r), canonical order: far-end sites of r owned by q, then all
    # operators on q's bonds that touch a K-site of r
    entries = {}
    for q in range(world):
        for r in range(world):
            if q == r:
                continue
            r_own = [s for s in range(N) if owner_site(s) == r]
            r_k = set(r_own) | {int(dst[b]) for b in range(len(src)) if owner_bond(b) == r}
            sites = sorted(s for s in r_k if owner_site(s) == q)
            bonds = sorted(b for b in range(len(src)) if owner_bond(b) == q and (int(src[b]) in r_k or int(dst[b]) in r_k))
            if rank not in (q, r):
                continue
            seg = [node_site[s] for s in sites]
            for b in bonds:
                seg += [op_node[k] for k, o in enumerate(ops) if (int(o["loc"]) >> 1) == b]
            entries[(q, r)] = seg
    return parent, length, entries


def worker(rank, world, port, lat, ops, ref_nc, ref_ssus, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parent, length, entries = local_forest(lat, ops, rank, world)
    roots = [find(parent, x) for x in range(len(parent))]
    csum = {}
    for x, r in enumerate(roots):
        csum[r] = csum.get(r, 0.0) + length[x]
    # my buffer: my segments in canonical order; every entry -> smallest entry of its local cluster
    keys = sorted(entries)
    flat, off = [], {}
    for kx in keys:
        off[kx] = len(flat)
        flat += entries[kx]
    rep = {}
    for i, x in enumerate(flat):
        rep.setdefault(roots[x], i)
    send = torch.tensor([rep[roots[x]] for x in flat], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(flat)], dtype=torch.int64))
    stride = int(max(s.item() for s in sizes))
    pad = torch.full((stride,), -1, dtype=torch.int64)
    pad[: len(flat)] = send
    recv = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(recv, pad)
    # every rank: forest over (rank, entry); with two ranks both segments involve both ranks, so both
    # sides know both layouts (the engine derives the layouts of ALL ranks from the static plan)
    gp = list(range(world * stride))
    for r in range(world):
        for i in range(stride):
            v = int(recv[r][i])
            if v >= 0:
                gp[r * stride + i] = r * stride + v
    other = 1 - rank
    for (qq, rr), seg in entries.items():
        # offsets of the same segment in the other rank's buffer: same canonical order of keys
        o_me, o_other = off[(qq, rr)], off[(qq, rr)]
        for i in range(len(seg)):
            union(gp, rank * stride + o_me + i, other * stride + o_other + i)
    used = [r * stride + i for r in range(world) for i in range(stride) if int(recv[r][i]) >= 0]
    groots = sorted({find(gp, x) for x in used})
    gid = {g: i for i, g in enumerate(groots)}
    table = torch.zeros(len(groots), dtype=torch.float64)
    open_roots = {roots[x] for x in flat}
    for r in open_roots:
        table[gid[find(gp, rank * stride + rep[r])]] += csum[r]
    dist.all_reduce(table)
    # closed clusters: roots that are owned nodes or own sites (ghost nodes nobody refers to are junk)
    N = lat["num_sites"]
    touched = set()
    src, dst = lat["src"], lat["dst"]
    for x, r in enumerate(roots):
        if length[x] != 0.0:
            touched.add(r)
    closed = [csum[r] for r in touched if r not in open_roots]
    nown = sum(1 for o in ops if int(src[int(o["loc"]) >> 1]) * world // N == rank)
    loc = torch.tensor([float(len(closed)), float(sum(v * v for v in closed)), float(nown)], dtype=torch.float64)
    dist.all_reduce(loc)
    nc = int(loc[0]) + len(groots)
    ssus = float(loc[1]) + float((table * table).sum())
    ok = (nc == ref_nc) and abs(ssus - ref_ssus) < 1e-9 * max(1.0, ref_ssus) and int(loc[2]) == len(ops)
    q.put((rank, ok, nc, ssus))
    dist.destroy_process_group()


@pytest.mark.parametrize("L,beta", [(16, 10.0), (12, 4.0)])
def test_two_rank_space_merge_over_gloo(L, beta):
    lat = ll.chain_lattice(L)
    sim = orc.OracleSim(lat, beta)
    for _ in range(150):
        sim.sweep()
    spins, ops = sim.get_state()
    _, ref_nc, ref = orc.build_clusters(lat, spins, ops)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=worker, args=(r, 2, port, lat, ops, ref_nc, ref["sa_ssus"], q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, nc, ssus in res:
        assert ok, (rank, nc, ref_nc, ssus, ref["sa_ssus"])
