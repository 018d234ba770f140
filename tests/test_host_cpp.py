"""Host-side C++ mirror of the reference API (alps-looper_b200/looper/): CPU unit test binary,
its union-find partition against the oracle's restatement of looper/union_find.h, and (GPU) the
loop driver against exact diagonalisation."""
import ctypes as C
import os
import re
import subprocess

import pytest

import oracle_util as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_host_test(tmp_path):
    exe = os.path.join(str(tmp_path), "test_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/cpp/test_host.cpp")])
    return exe


def test_host_mirror_unit(tmp_path):
    exe = _build_host_test(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "host ok" in out.stdout
    # min-index roots of the host union_find == partition of the reference union_find on the
    # same 100 unions (test/union_find.op, via the oracle replay text)
    roots = [int(x) for x in out.stdout.splitlines()[0].split()[1:]]
    n = orc.lib().orc_union_find_replay(None, 0)
    buf = C.create_string_buffer(n + 1)
    orc.lib().orc_union_find_replay(buf, n + 1)
    ref_root = list(range(100))
    for m in re.finditer(r"node (\d+)'s parent is \d+ and its root is (\d+)", buf.value.decode().split("[results]")[1]):
        ref_root[int(m.group(1))] = int(m.group(2))
    groups = {}
    for i, r in enumerate(ref_root[:100]):
        groups.setdefault(r, []).append(i)
    canon = list(range(100))
    for members in groups.values():
        for i in members:
            canon[i] = min(members)
    assert roots == canon


@pytest.mark.gpu
def test_loop_driver_vs_exact_diagonalisation():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "alps-looper_b200/looper")])
    exe = os.path.join(ROOT, "alps-looper_b200/looper/loop")
    out = subprocess.run([exe, "-l", "8", "-t", "0.2", "-n", "16384"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = {}
    for ln in out.stdout.splitlines():
        m = re.match(r"(.+?)\s*=\s*(\S+) \+- (\S+)", ln)
        if m:
            vals[m.group(1).strip()] = (float(m.group(2)), float(m.group(3)))
    ed = {"Energy Density": -0.441438, "Uniform Susceptibility": 0.0804441,
          "Staggered Magnetization^2": 6.59939, "Staggered Susceptibility": 2.40159}
    for k, ex in ed.items():
        mean, err = vals[k]
        assert abs(mean - ex) < 4 * err + 1e-9, (k, mean, err, ex)


@pytest.mark.gpu
def test_loop_driver_transverse_field_vs_exact_diagonalisation():
    """Gamma (site graphs) and MEASURE[Stiffness] through the host worker, against
    tests/golden/ed_tfi.json row 2 (Jxy = -1, Jz = 0.5, Gamma = 0.6, L = 8, T = 0.4)."""
    import json
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "alps-looper_b200/looper")])
    exe = os.path.join(ROOT, "alps-looper_b200/looper/loop")
    ed = json.load(open(os.path.join(ROOT, "tests/golden/ed_tfi.json")))[2]
    params = ('LATTICE = "chain lattice"; L = 8; Jxy = -1; Jz = 0.5; Gamma = 0.6; T = 0.4; '
              'SWEEPS = 32768; VERBOSE = 1; MEASURE[Stiffness] = 1;\n')
    out = subprocess.run([exe, "-"], input=params, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = {}
    for ln in out.stdout.splitlines():
        m = re.match(r"(.+?): (\S+) \+/- (\S+);", ln)
        if m:
            vals[m.group(1).strip()] = (float(m.group(2)), float(m.group(3)))
    for name, key in (("Energy Density", "energy_density"), ("Transverse Magnetization Density", "transmag_density")):
        mean, err = vals[name]
        assert abs(mean - ed[key]) < 5 * err + 1e-9, (name, mean, ed[key], err)
    assert vals["Stiffness"][0] >= 0


@pytest.mark.gpu
def test_loop_driver_checkpoint_resume(tmp_path):
    """save()/load() of the worker (path_integral.C:111-124 field order) through the driver: the second
    run resumes where the first stopped and its observables are those of the same ensemble."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "alps-looper_b200/looper")])
    exe = os.path.join(ROOT, "alps-looper_b200/looper/loop")
    ck = os.path.join(str(tmp_path), "run.ck")
    first = subprocess.run([exe, "-l", "8", "-t", "0.2", "-n", "8192", "--checkpoint", ck], capture_output=True, text=True)
    assert first.returncode == 0, first.stderr
    assert os.path.getsize(ck) > 8 + 4 + 4 + 8 * 4 + 8
    assert open(ck, "rb").read(8) == b"LQCKPT02"
    second = subprocess.run([exe, "-l", "8", "-t", "0.2", "-n", "24576", "--checkpoint", ck], capture_output=True, text=True)
    assert second.returncode == 0, second.stderr
    assert "resumed at" in second.stdout
    frac = float(re.search(r"resumed at (\S+)", second.stdout).group(1))
    assert 0.3 < frac < 0.4            # (1024 + 8192) of (3072 + 24576) steps
    m = re.search(r"Energy Density\s*=\s*(\S+) \+- (\S+)", second.stdout)
    mean, err = float(m.group(1)), float(m.group(2))
    assert abs(mean + 0.441438) < 5 * err + 1e-6
    bad = subprocess.run([exe, "-l", "10", "-t", "0.2", "-n", "1024", "--checkpoint", ck], capture_output=True, text=True)
    assert bad.returncode != 0 and ("lattice size differs" in bad.stderr or "tiling differ" in bad.stderr)
    # a finished run resumed again measures nothing more and says so instead of printing nan
    third = subprocess.run([exe, "-l", "8", "-t", "0.2", "-n", "24576", "--checkpoint", ck], capture_output=True, text=True)
    assert third.returncode == 0 and "nan" not in third.stdout
    m3 = re.search(r"Energy Density\s*=\s*(\S+) \+- (\S+)", third.stdout)
    assert m3 and float(m3.group(1)) == pytest.approx(mean, rel=1e-12)      # the observables travel with the checkpoint


@pytest.mark.gpu
def test_loop_driver_sse_algorithm(tmp_path):
    """ALGORITHM = "loop; sse" (sse.C:411) through the host mirror: the SSE forms of the estimators
    (susceptibility.h:213-215) against exact diagonalisation (chain L = 8, T = 0.2)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "alps-looper_b200/looper")])
    exe = os.path.join(ROOT, "alps-looper_b200/looper/loop")
    par = os.path.join(str(tmp_path), "sse.ip")
    open(par, "w").write('ALGORITHM = "loop; sse"\nLATTICE = "chain lattice"\nL = 8\nT = 0.2\nSWEEPS = 32768\n')
    out = subprocess.run([exe, par], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = {}
    for ln in out.stdout.splitlines():
        m = re.match(r"(.+?)\s*=\s*(\S+) \+- (\S+)", ln)
        if m:
            vals[m.group(1).strip()] = (float(m.group(2)), float(m.group(3)))
    for name, ex in (("Energy Density", -0.441438), ("Staggered Susceptibility", 2.40159), ("Staggered Magnetization^2", 6.59939)):
        mean, err = vals[name]
        assert abs(mean - ex) < 5 * err + 1e-9, (name, mean, ex, err)
    bad = subprocess.run([exe, "-"], input='ALGORITHM = "loop; worm"\n', capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown ALGORITHM" in bad.stderr
