#!/usr/bin/env python
"""bench.py -- loop-update throughput (operators/s and MCS/s) of the B200 engine.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line.  A "step" is one Monte Carlo step (diagonal update -> union-find labelling -> estimators ->
cluster flip) of the S=1/2 Heisenberg antiferromagnet on the square lattice named in
BASELINE.json.  `value` = operators processed per second, whole job, state resident in HBM,
device-timed with CUDA events on the engine's stream.  `e2e` = the same metric through the
reference-facing call (lq_sweep == loop_worker::run): one call per step, per-step inputs copied
from pinned host memory, the step's collector read back to the host, host clock.

`--impl reference` times the reference's own CPU algorithm (the oracle port of standalone/loop.C,
validated bit-for-bit against the reference binary; the reference binary itself only knows the
chain lattice) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALGO_BYTES_PER_OP = 104.0
# algorithmic bytes per operator of the individual phases
# (the operator flip, 12 B, is fused into k_estimate: phase 12 carries 32 + 12; phase 15 is the spin flip)
PHASE_BYTES = {5: ("k_diag_update", 24.0), 7: ("k_walk+k_union_local+k_union_global", 28.0),
               11: ("k_compress+k_relabel", 8.0), 12: ("k_estimate", 44.0)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def workload(name):
    """(lattice extents, beta, thermalisation sweeps, SSE representation) of the named synthetic workload."""
    table = {
        "square1024_beta1024": ((1024, 1024), 1024.0, 64, False),   # stationarity trace: profiles/r02_thermalisation_trace.md
        "square1024_beta128": ((1024, 1024), 128.0, 40, False),
        "square256_beta64": ((256, 256), 64.0, 200, False),
        "square64_beta8": ((64, 64), 8.0, 100, False),
        # BASELINE config 4: simple cubic 64^3 near T_N = 0.946 J, SSE representation (sse.C)
        "cubic64_T0.95_sse": ((64, 64, 64), 1.0 / 0.95, 400, True),
        "cubic64_T0.95": ((64, 64, 64), 1.0 / 0.95, 400, False),
    }
    return table[name]


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample, one independent replica per core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    L, beta, therm, steps, seed = args
    import numpy as np  # noqa: F401
    import oracle_util as orc
    n = L * L
    import numpy as np
    idx = np.arange(n)
    x, y = idx % L, idx // L
    src = np.concatenate([idx, idx]).astype(np.int32)
    dst = np.concatenate([(x + 1) % L + L * y, x + L * ((y + 1) % L)]).astype(np.int32)
    lat = dict(num_sites=n, src=src, dst=dst, gauge=np.where((x + y) % 2 == 0, 1.0, -1.0))
    sim = orc.OracleSim(lat, beta, seed, looper_estimators=False)   # standalone/loop.C statements only
    for _ in range(therm):
        sim.sweep()
    times, nops = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        c = sim.sweep()
        times.append(time.perf_counter() - t0)
        nops.append(c["nop"])
    return times, nops


def cpu_arm(L, beta, therm, steps, cores):
    import oracle_util as orc
    orc.build()
    jobs = [(L, beta, therm, steps, 29833 + 7919 * i) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    # aggregate throughput of the replicas over the timed steps
    per_core = [sum(n) / sum(t) for t, n in res]
    ms_step = 1e3 * sum(sum(t) for t, _ in res) / (len(res) * steps)
    nop = sum(sum(n) for _, n in res) / (len(res) * steps)
    return dict(ops_per_s=sum(per_core), ms_per_step=ms_step, nop=nop, wall=wall)


def reference_binary_check():
    """The unmodified reference binary (oracle/_ref/loop = standalone/loop.C, chains only) and the
    oracle port on the same chain: operators/s of both, to show that the port used for the 2-D
    workload is a fair stand-in for the reference's own code.  None if the binary is absent."""
    import re
    ref = os.path.join(ROOT, "oracle", "_ref", "loop")
    port = os.path.join(ROOT, "oracle", "oracle_loop")
    if not (os.path.exists(ref) and os.path.exists(port)):
        return None
    L, T, n = 16384, 0.0625, 48
    out = {"workload": f"chain L={L} T={T}, {n} MCS + thermalisation, single thread"}
    try:
        for name, exe in (("reference", ref), ("port", port)):
            t0 = time.perf_counter()
            txt = subprocess.run([exe, "-l", str(L), "-t", str(T), "-n", str(n)], capture_output=True, text=True,
                                 timeout=120).stdout
            wall = time.perf_counter() - t0
            ene = float(re.search(r"Energy Density\s*=\s*(\S+)", txt).group(1))
            nop = (0.25 - ene) * L / T      # standalone/loop.C:175: E = (B/4 - n/beta) / L
            out[name + "_operators_per_s"] = nop * (n + (n >> 3)) / wall   # process wall clock, all MCS
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}
    return out


def reference_parallel_check(cores):
    """The reference's own parallel code: standalone/loop_mpi.C, unmodified, one rank per core over
    the thread-backed mpi.h of oracle/mpi_shim (oracle/_ref/loop_mpi) on a long chain; the serial
    reference binary on the same chain beside it.  None if the binaries are absent."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "loop_mpi")
    ser = os.path.join(ROOT, "oracle", "_ref", "loop")
    if not (os.path.exists(exe) and os.path.exists(ser)):
        return None
    # standalone/parallel.h:240 sends 2N estimates out of a vector that holds fewer (reads past its end; harmless
    # under a real MPI, a segmentation fault here when the copy crosses an unmapped page), so the run is retried
    # on shorter chains until one completes
    T, n = 1.0 / 16, 12
    out = {"ranks": cores}
    for L in (1 << 20, 1 << 18, 1 << 16, 1 << 14):
        try:
            res = {}
            for name, cmd in (("serial", [ser]), ("parallel", [exe, str(cores)])):
                t0 = time.perf_counter()
                txt = subprocess.run(cmd + ["-l", str(L), "-t", str(T), "-n", str(n)], capture_output=True, text=True,
                                     timeout=300).stdout
                wall = time.perf_counter() - t0
                ene = float(re.search(r"Energy Density\s*=\s*(\S+)", txt).group(1))
                nop = (0.25 - ene) * L / T
                res[name + "_operators_per_s"] = nop * (n + (n >> 3)) / wall
            out.update(res)
            out["workload"] = f"chain L={L} T={T}, {n} MCS + {n >> 3} thermalisation (process wall clock)"
            return out
        except Exception as e:  # noqa: BLE001
            out["note"] = "standalone/loop_mpi.C over the thread shim failed on longer chains (parallel.h:240 over-read)"
            continue
    out["error"] = "no chain length completed"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 64))
    # out-of-cache sample of the same model / lattice family: 512 x 512, beta = 16 is 5e6 operators
    # (~0.4 GB per replica); thermalised from the empty state like the GPU arm
    L, beta, therm = 512, 16.0, 12
    r = cpu_arm(L, beta, therm, args.steps + args.warmup, cores)
    sample = (f"oracle port of standalone/loop.C (its statements only; validated bit-for-bit against loop.op and "
              f"the reference binary), {cores} independent replicas of square {L}x{L} beta={beta:g} "
              f"(same model/lattice family as the GPU workload, out of cache, bounded size), "
              f"{therm} thermalisation + {args.warmup + args.steps} timed MCS each")
    line = {
        "impl": "reference", "metric": "loop_update_operators_per_sec", "value": r["ops_per_s"],
        "unit": "operators/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": f"square{L}_beta{beta:g}",
                   "operators_per_mcs": r["nop"]},
        "cpu_baseline": {"value": r["ops_per_s"], "unit": "operators/s", "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": r["ops_per_s"], "unit": "operators/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    chk = reference_binary_check()
    if chk is not None:
        line["reference_binary_check"] = chk
    par = reference_parallel_check(cores)
    if par is not None:
        line["reference_parallel_check"] = par
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import looper_b200 as lq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dims, beta, therm, sse = workload(args.workload)
    L = dims[0]
    if args.therm >= 0:
        therm = args.therm
    lat = lq.hypercubic_lattice(dims)
    tile = args.tile_sites or (256 if lat["num_sites"] >= 512 * 512 else 64)
    args.sse = sse
    # parity pre-flight (not timed; the oracle is the checker here, never the thing measured): one
    # oracle configuration through the same engine / communicator, against the reference union-find
    parity = parity_preflight(rank, world, local, args)
    eng = make_engine(lq, lat, beta, tile, local, rank, world, args)
    info = eng.info()
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # thermalise (not timed) + warm-up steps; the clock sampler runs from here on (nvidia-smi needs
    # ~1 s to start), every sample is taken under load
    sampler = ClockSampler(local)
    eng.sweep_many(therm, collect=False)
    eng.sweep_many(max(args.warmup, 3), collect=False)

    # ---- device-timed region: K steps, state resident in HBM -----------------------------------
    launches0 = eng.kernel_launches()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        out = eng.sweep_many(args.steps)
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.kernel_launches() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    nops = torch.tensor([float(out["nop"].sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # (every rank reports the global operator count of the shared configuration)
    ms = float(t.item())
    total_ops = float(nops.item())
    value = total_ops / (ms * 1e-3)

    # ---- end-to-end through the reference-facing call (lq_sweep == loop_worker::run) -----------
    h0, d0 = eng.copied_bytes()
    barrier()
    t0 = time.perf_counter()
    e2e_ops = 0.0
    for _ in range(args.steps):
        c = eng.sweep()          # pinned H2D of the step inputs, kernels, D2H of the collector, sync
        e2e_ops += c["nop"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h1, d1 = eng.copied_bytes()
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ne = torch.tensor([e2e_ops], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(ne.item()) / float(te.item())

    # ---- per-kernel durations (CUDA events around each phase, separate short run) --------------
    roof = None
    if True:  # every rank runs these steps (they contain collectives); rank 0 reports
        peak, which = measured_peaks()
        eng_t = eng
        if eng_t is not None:
            eng_t.enable_timers(True)
            before = {x["id"]: (x["seconds"], x["count"]) for x in eng_t.timers()}
            nt = max(3, min(args.steps, 10))
            o2 = eng_t.sweep_many(nt)
            after = {x["id"]: (x["seconds"], x["count"]) for x in eng_t.timers()}
            nop_t = float(o2["nop"].mean())
            phases = {}
            for pid, (sec, cnt) in after.items():
                s0, c0 = before.get(pid, (0.0, 0))
                if cnt > c0:
                    phases[pid] = (sec - s0) / nt   # per STEP (a multi-rank step enters phases 8 and 13 more than once)
            tot = sum(phases.values())
            dom = max((p for p in phases if p in PHASE_BYTES), key=lambda p: phases[p])
            kname, bpo = PHASE_BYTES[dom]
            traffic = None
            try:  # dram__bytes_read+write per launch of that kernel from the committed ncu capture
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                if tj.get("workload") == args.workload and world == 1:
                    traffic = sum(tj["kernels"][k]["dram_bytes_per_launch"] for k in kname.split("+"))
            except Exception:
                traffic = None
            ach = (nop_t / world) * bpo / phases[dom] / 1e9   # this rank's slab
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": which,
                    "kernel_ms": 1e3 * phases[dom], "kernel_share_of_step": phases[dom] / tot,
                    "algorithmic_bytes_per_op": bpo,
                    "phase_ms": {str(k): 1e3 * v for k, v in sorted(phases.items())},
                    "step_frac": value / world * ALGO_BYTES_PER_OP / (peak * 1e9),
                    "step_bytes_per_op": ALGO_BYTES_PER_OP}
            eng_t.enable_timers(False)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_arm(1024, 16.0, 8, 5, 1)
        cpu = {"value": r["ops_per_s"], "unit": "operators/s", "cores": 1, "kind": "port",
               "sample": "oracle port of standalone/loop.C (its statements only, single-threaded like the reference), "
                         "square 1024x1024 beta=16 (the headline lattice at reduced beta: out of cache), "
                         "8 thermalisation + 5 timed MCS, %.0f operators/MCS" % r["nop"]}

    if rank == 0:
        nop_mean = float(out["nop"].mean())
        line = {
            "metric": "loop_update_operators_per_sec", "value": value, "unit": "operators/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong",   # ONE Markov chain of fixed size, split over the GPUs
            "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "mcs_per_sec": args.steps / (ms * 1e-3),
            "config": {"workload": args.workload, "lattice": ("square " if len(dims) == 2 else "simple cubic ") +
                       "x".join(str(x) for x in dims) + " periodic",
                       "model": "S=1/2 Heisenberg AF J=1", "beta": beta,
                       "representation": "sse (sse.C estimators, string positions from a counting sort)" if sse else "path integral",
                       "operators_per_mcs": nop_mean, "clusters_per_mcs": float(out["nc"].mean()),
                       "open_clusters_per_mcs": float(out["noc"].mean()),
                       "thermalisation_mcs": therm, "tile_sites": tile, "windows": info["num_windows"],
                       "page_capacity": info["page_capacity"], "reserve": args.reserve,
                       "arena_regrows": eng.regrow_count(),
                       "tiles": info["num_tiles"], "device_bytes": info["device_bytes"],
                       "l2_policy": "working set (%.1f GB) larger than L2" % (info["device_bytes"] / 1e9)
                       if info["device_bytes"] > 2.5e8 else "working set comparable to L2",
                       "multi_gpu": multi_gpu_mode(args, world)},
            "e2e": {"value": e2e_value, "unit": "operators/s",
                    "h2d_bytes_per_step": (h1 - h0) / args.steps,
                    "d2h_bytes_per_step": (d1 - d0) / args.steps,
                    "ms_per_step": 1e3 * float(te.item()) / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line["parity"] = parity
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_preflight(rank, world, local, args):
    try:
        if world > 1:
            import mgpu_parity
            return mgpu_parity.preflight(rank, world, local, args.comm, steps=6, cut=args.cut)
        import numpy as np
        import looper_b200 as lq
        import oracle_util as orc
        lat = lq.hypercubic_lattice((64, 64))
        sim = orc.OracleSim(lat, 8.0, 29833)
        for _ in range(40):
            sim.sweep()
        spins, ops = sim.get_state()
        ref_labels, ref_nc, _ = orc.build_clusters(lat, spins, ops)
        eng = lq.Engine(lat, 8.0, device=local, tile_sites=64)
        eng.set_state(spins, ops)
        labels, nc, _ = eng.build_clusters()
        eng.close()
        ok = bool(nc == ref_nc and np.array_equal(labels, ref_labels))
        return {"ok": ok, "ranks": 1, "cases": [{"name": "square64_b8 partition bit-exact vs oracle", "ok": ok,
                                                  "operators": int(len(ops)), "nc": int(nc)}]}
    except Exception as e:  # noqa: BLE001
        return {"ok": False, "error": repr(e)[:300]}


def multi_gpu_mode(args, world):
    if world == 1:
        return "single GPU"
    if args.cut == "space":
        return ("spatial strips: %d ranks, every rank owns a contiguous range of tiles over the whole imaginary-time "
                "axis + ghost copies of the neighbouring tile rows (halo pages and spins by ncclSend/ncclRecv twice per "
                "step), boundary segments all-gathered and open-cluster sums all-reduced with NCCL every step (one "
                "Markov chain over all GPUs); issued by %s"
                % (world, "the engine (lq_comm_init)" if args.comm == "nccl" else "torch.distributed callbacks (lq_set_comm)"))
    return ("imaginary-time slabs: %d ranks, boundary cluster ids all-gathered and open-cluster sums "
            "all-reduced with NCCL every step (one Markov chain over all GPUs); collectives issued by %s"
            % (world, "the engine (lq_comm_init, ncclAllGather/ncclAllReduce on its stream)" if args.comm == "nccl"
               else "torch.distributed callbacks (lq_set_comm)"))


def load_comm():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lq_comm", os.path.join(ROOT, "alps-looper_b200", "comm.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_engine(lq, lat, beta, tile, local, rank, world, args, timers=False):
    eng = lq.Engine(lat, beta, seed=29833, device=local, tile_sites=tile, timers=timers,
                    window_ops=args.window_ops, reserve=args.reserve,
                    rank=rank if world > 1 else 0, nranks=world, sse=getattr(args, "sse", False),
                    cut=args.cut if world > 1 else "time")
    if world > 1:
        if args.comm == "nccl":
            load_comm().attach_nccl(eng, rank, world)
        else:
            load_comm().attach_torch_distributed(eng, local)
    return eng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LQ_BENCH_WORKLOAD", "square1024_beta1024"))
    ap.add_argument("--tile-sites", type=int, default=0)
    ap.add_argument("--window-ops", type=float, default=0.0, help="candidates per bond and window (0 = engine default)")
    ap.add_argument("--reserve", type=float, default=1.4,
                    help="page capacity / mean candidate count (engine default 1.7; the Heisenberg workloads "
                         "hold 1.17 operators per candidate)")
    ap.add_argument("--therm", type=int, default=-1, help="override thermalisation sweeps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--comm", default="nccl", choices=["nccl", "torch"],
                    help="multi-GPU data plane: the engine's own NCCL communicator (lq_comm_init) or "
                         "torch.distributed callbacks (lq_set_comm)")
    ap.add_argument("--cut", default=os.environ.get("LQ_BENCH_CUT", "time"), choices=["time", "space"],
                    help="multi-GPU decomposition of the one Markov chain: imaginary-time slabs (looper/parallel.h) "
                         "or spatial strips (BASELINE config 3 wording; lq_options.cut)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
