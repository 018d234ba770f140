"""Real multi-GPU run of the slab engine over torch.distributed/NCCL (skipped on a 1-GPU box;
the same exchange logic is covered on one GPU by tests/test_gpu_slabs.py and on CPU by
tests/test_slab_protocol_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_bench_line():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "4", "--warmup", "3", "--workload", "square256_beta64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0
    assert 4.0e6 < line["config"]["operators_per_mcs"] < 6.0e6


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("comm", ["nccl", "torch"])
def test_slab_parity_over_real_nccl(comm):
    """tests/mgpu_parity.py on every GPU of the box (up to 8): merged cluster count and sums of an
    oracle configuration equal the oracle's, slab steps stay legal, collectors identical."""
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29541" if comm == "nccl" else "29542",
           os.path.join(ROOT, "tests", "mgpu_parity.py"), "--comm", comm]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MGPU_PARITY ")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-2000:])
    rep = json.loads(lines[-1][len("MGPU_PARITY "):])
    assert rep["ok"] and rep["ranks"] == n and all(c["ok"] for c in rep["cases"]), rep


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("comm", ["nccl", "torch"])
def test_space_parity_over_real_nccl(comm):
    """the same check for the SPATIAL cut (lq_options.cut = LQ_CUT_SPACE): halo pages and spins travel
    by ncclSend/ncclRecv (or torch.distributed P2P), boundary segments through the all-gather."""
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29543" if comm == "nccl" else "29544",
           os.path.join(ROOT, "tests", "mgpu_parity.py"), "--comm", comm, "--cut", "space"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MGPU_PARITY ")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-2000:])
    rep = json.loads(lines[-1][len("MGPU_PARITY "):])
    assert rep["ok"] and rep["ranks"] == n and rep["cut"] == "space" and all(c["ok"] for c in rep["cases"]), rep


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("partition", ["time", "space"])
def test_cpp_driver_two_ranks_vs_exact_diagonalisation(partition):
    """the C++ host mirror (looper::loop_worker with a communicator, path_integral_mpi.C:75) over two GPUs,
    no Python in the run: `loop --nranks 2` forks one process per GPU, the NCCL id travels through a file,
    PARTITION picks the cut; observables against exact diagonalisation of the L = 8 chain at T = 0.2."""
    import re
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "alps-looper_b200/looper")])
    exe = os.path.join(ROOT, "alps-looper_b200/looper/loop")
    params = ('LATTICE = "chain lattice"; L = 8; T = 0.2; SWEEPS = 16384; TILE_SITES = 2; '
              'PARTITION = "%s";\n' % partition)
    out = subprocess.run([exe, "--nranks", "2", "-"], input=params, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    vals = {}
    for ln in out.stdout.splitlines():
        m = re.match(r"(.+?)\s*=\s*(\S+) \+- (\S+)", ln)
        if m:
            vals[m.group(1).strip()] = (float(m.group(2)), float(m.group(3)))
    ed = {"Energy Density": -0.441438, "Staggered Magnetization^2": 6.59939, "Staggered Susceptibility": 2.40159}
    for k, ex in ed.items():
        mean, err = vals[k]
        assert abs(mean - ex) < 4 * err + 1e-9, (k, mean, err, ex)
