"""Multi-GPU parity of the imaginary-time slab engine over REAL NCCL, one process per GPU.

Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
          --master-port 29541 tests/mgpu_parity.py [--comm nccl|torch]
(tests/test_gpu_multi.py launches it when the box has >= 2 GPUs; bench.py --gpus N runs `preflight`
before its timed region and carries the verdict in its JSON line.)

Every rank thermalises the SAME configuration with the oracle (reference CPU algorithm, same seed),
loads it into its slab engine (rank r owns tau in [r/P, (r+1)/P)), and the merged result of
lq_build_clusters -- number of clusters and the 14 cluster sums, after the all-gather of the
boundary ids and the all-reduce of the open-cluster sums (looper/parallel.h:1609-1809) -- must equal
the oracle's union-find on the whole configuration (standalone/loop_mpi.C:55-68 is the reference's
own version of this check: the P-rank run reproduces the serial numbers).  Then the slab engines
run Monte Carlo steps; the union of their slabs must stay a legal world-line configuration and
every rank must report the same collector.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SUMS = ["umag0", "usize2", "umag2", "usize4", "umag4", "usize", "umag",
        "smag0", "ssize2", "smag2", "ssize4", "smag4", "ssize", "smag"]


def _comm_mod():
    spec = importlib.util.spec_from_file_location("lq_comm", os.path.join(ROOT, "alps-looper_b200", "comm.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def attach(eng, local, rank, world, comm):
    m = _comm_mod()
    if comm == "torch":
        m.attach_torch_distributed(eng, local)
    else:
        m.attach_nccl(eng, rank, world)


def preflight(rank, world, local, comm="nccl", cases=None, steps=12, cut="time"):
    """returns a dict (identical on all ranks): {"ok": bool, "cases": [...]}; raises nothing."""
    import torch
    import torch.distributed as dist
    import looper_b200 as lq
    import oracle_util as orc

    if cases is None:
        # (spatial cut: at least one tile per rank)
        cases = [("chain64_b16", lq.chain_lattice(64), 16.0, 120, 4 if cut == "space" else 0),
                 ("square32_b8", lq.hypercubic_lattice((32, 32)), 8.0, 60, 16)]
    report = {"ok": True, "ranks": world, "comm": comm, "cut": cut, "cases": []}
    for name, lat, beta, therm, tile in cases:
        rec = {"name": name}
        try:
            sim = orc.OracleSim(lat, beta, 29833)
            for _ in range(therm):
                sim.sweep()
            spins, ops = sim.get_state()
            _, ref_nc, ref = orc.build_clusters(lat, spins, ops)
            eng = lq.Engine(lat, beta, seed=99, device=local, tile_sites=tile, rank=rank, nranks=world, cut=cut)
            attach(eng, local, rank, world, comm)
            eng.set_state(spins, ops)
            nloc = eng.num_ops()
            nc = C.c_int64(0)
            c = lq.LqCollector()
            lq._check(lq.lib.lq_build_clusters(eng._h, None, C.byref(nc), C.byref(c)))
            d = c.as_dict()
            tot = torch.tensor([nloc], dtype=torch.int64, device="cuda")
            dist.all_reduce(tot)
            good = (nc.value == ref_nc) and (int(tot.item()) == len(ops)) and d["nop"] == len(ops)
            worst = 0.0
            for f in SUMS:
                den = max(abs(ref[f]), 1e-7)
                worst = max(worst, abs(d[f] - ref[f]) / den)
            good = good and worst < 1e-8
            rec.update(nc=int(nc.value), ref_nc=int(ref_nc), operators=int(len(ops)),
                       open_clusters=float(d["noc"]), max_rel_diff_sums=worst)
            # Monte Carlo steps of the slab engines: identical collectors, legal union of the slabs
            out = eng.sweep_many(steps)
            s2, o2 = eng.get_state()
            sig = torch.tensor([float(out["nop"].sum()), float(out["nc"].sum()), float(out["usize"].sum())],
                               dtype=torch.float64, device="cuda")
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo, hi))
            gathered = [None] * world
            dist.all_gather_object(gathered, (o2.tobytes(), s2.tobytes() if (rank == 0 or cut == "space") else b""))
            legal = True
            if rank == 0:
                allops = np.concatenate([np.frombuffer(g[0], dtype=lq.OP_DTYPE) for g in gathered])
                order = np.argsort(allops["time"], kind="stable")
                if cut == "space":   # every rank reports the spins of its own sites, -1 elsewhere
                    s2 = np.full(lat["num_sites"], -1, dtype=np.int32)
                    for g in gathered:
                        sg = np.frombuffer(g[1], dtype=np.int32)
                        s2[sg >= 0] = sg[sg >= 0]
                try:
                    orc.build_clusters(lat, s2, allops[order])
                    legal = len(allops) == int(out["nop"][-1])
                except Exception:   # noqa: BLE001
                    legal = False
            flag = torch.tensor([1 if (good and same and legal) else 0], dtype=torch.int64, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            rec.update(collectors_identical=same, union_legal=legal, ok=bool(flag.item()))
            eng.close()
        except Exception as e:   # noqa: BLE001
            rec.update(ok=False, error=repr(e)[:300])
        report["cases"].append(rec)
        report["ok"] = report["ok"] and rec["ok"]
    return report


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="nccl", choices=["nccl", "torch"])
    ap.add_argument("--cut", default="time", choices=["time", "space"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rep = preflight(rank, world, local, args.comm, cut=args.cut)
    if rank == 0:
        print("MGPU_PARITY " + json.dumps(rep), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if rep["ok"] else 1)


if __name__ == "__main__":
    main()
