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


def _loop_exe():
    exe = os.path.join(ROOT, "alps-looper_b200", "looper", "loop")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)], stdout=subprocess.DEVNULL)
    return exe


def test_loop_driver_reads_alps_parameter_files_dry_run(tmp_path):
    """`loop --dry-run` (no GPU): an ALPS-style parameter file -- globals, tasks in braces, expressions,
    several statements per line -- becomes one validated task per block; what the accelerated path does not
    implement is refused per task with a reason, the other tasks stand."""
    f = tmp_path / "params"
    f.write_text('LATTICE = "chain lattice"\nMODEL = "spin"\nlocal_S = 1/2; L = 4; Jxy = -1; Jz = -1\n'
                 'ALGORITHM = "loop; path integral"\nSWEEPS = 2*512\n'
                 '{ T = 0.1 } { T = 1/L; ALGORITHM = "loop; sse" }\n{ T = 0.5; local_S = 1 }\n'
                 'LATTICE = "site"; Gamma = 0.7; Jxy = 0; Jz = 0\n{ T = 0.2; MEASURE[Correlations] = true }\n'
                 '{ T = 0.2; h = 0.3 }\n{ ALGORITHM = "diagonalization" }\n')
    out = subprocess.run([_loop_exe(), "--dry-run", str(f)], capture_output=True, text=True)
    lines = out.stdout.splitlines()
    assert out.returncode == 1                                   # some tasks were refused
    assert [ln for ln in lines if ln.startswith("[task")] == [
        "[task 1 of 6] T = 0.1;", "[task 2 of 6] T = 1/L; ALGORITHM = loop; sse;", "[task 3 of 6] T = 0.5; local_S = 1;",
        "[task 4 of 6] T = 0.2; MEASURE[Correlations] = true;", "[task 5 of 6] T = 0.2; h = 0.3;",
        "[task 6 of 6] ALGORITHM = diagonalization;"]
    ok = [ln for ln in lines if ln.startswith("ok:")]
    assert ok == ["ok: loop; path integral, 4 sites, 4 bonds, T = 0.1, graph weight 2, 128 + 1024 sweeps",
                  "ok: loop; sse, 4 sites, 4 bonds, T = 0.25, graph weight 2, 128 + 1024 sweeps",
                  "ok: loop; path integral, 1 sites, 0 bonds, T = 0.2, graph weight 0.35, 128 + 1024 sweeps"]
    err = out.stderr
    assert "error in task 3: local_S != 1/2" in err and "error in task 5: longitudinal fields" in err
    assert "error in task 6: unknown ALGORITHM 'diagonalization'" in err
    assert "warning: MEASURE[Correlations] is not measured" in err
    # a single task keeps the old behaviour: the error is fatal and plain
    out = subprocess.run([_loop_exe(), "--dry-run", "-"], input='MODEL = "XYZ spin"; Jx = 1\n', capture_output=True, text=True)
    assert out.returncode == 1 and out.stderr.startswith("error: MODEL 'XYZ spin' is not supported")
    out = subprocess.run([_loop_exe(), "--dry-run", "-l", "16", "-t", "0.1", "-n", "64"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok: loop; path integral, 16 sites, 16 bonds, T = 0.1")


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree absent (GPU box)")
def test_loop_driver_dry_run_on_the_reference_input_files():
    """The reference's own parameter files (loop.ip, check/*, extras/*/*.ip) parse into the tasks ALPS would make
    of them: the S = 1/2 loop tasks on built-in lattices are accepted, everything else is refused with a reason --
    nothing is 'not readable', nothing crashes."""
    import glob
    files = [os.path.join(REF, "loop.ip")] + sorted(glob.glob(os.path.join(REF, "extras", "*", "*.ip"))) + \
        [p for p in sorted(glob.glob(os.path.join(REF, "check", "*-*"))) if not p.endswith((".in", ".sh"))]
    assert len(files) > 40
    accepted = 0
    for path in files:
        out = subprocess.run([_loop_exe(), "--dry-run", path], capture_output=True, text=True, timeout=60)
        assert out.returncode in (0, 1), (path, out.returncode, out.stderr[-300:])
        assert "not readable" not in out.stderr and "nested" not in out.stderr and "missing '}'" not in out.stderr, (path, out.stderr[-300:])
        ntask = sum(ln.startswith("[task") for ln in out.stdout.splitlines()) or 1
        nok = sum(ln.startswith("ok:") for ln in out.stdout.splitlines())
        nerr = out.stderr.count("error")
        assert nok + nerr == ntask, (path, ntask, nok, nerr)
        accepted += nok
    loop_ip = subprocess.run([_loop_exe(), "--dry-run", os.path.join(REF, "loop.ip")], capture_output=True, text=True)
    assert loop_ip.stdout.count("[task") == 21 and loop_ip.stdout.count("ok:") == 5    # chain PI + SSE at T = 1/4, Ising PI + SSE, Heisenberg PI
    assert accepted > 100


def _fake_engine(tmp_path):
    """tests/cpp/fake_lq.c: a TEST DOUBLE of the C ABI (canned collectors, no physics), preloaded in front of
    liblq.so so that the host plumbing runs end to end without a GPU.  Never part of the product."""
    so = os.path.join(str(tmp_path), "libfake_lq.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests/cpp/fake_lq.c")])
    return dict(os.environ, LD_PRELOAD=so)


def test_loop_driver_plumbing_end_to_end_with_a_fake_engine(tmp_path):
    """Driver + worker + observables + evaluators + checkpoint around a test double of the engine: several tasks
    from one ALPS-style file, the five lines of standalone/loop.C per task, the reference's evaluated observables in
    the VERBOSE dump, and a checkpoint that resumes where the first run stopped."""
    env = _fake_engine(tmp_path)
    f = tmp_path / "params"
    f.write_text('LATTICE = "chain lattice"; L = 4; Jxy = -1; SWEEPS = 256; VERBOSE = 1\n'
                 '{ T = 0.5 } { T = 1/L; Gamma = 0.4 }\n{ T = 0.5; local_S = 1 }\n')
    out = subprocess.run([_loop_exe(), str(f)], capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 1 and "error in task 3: local_S != 1/2" in out.stderr, out.stderr
    blocks = out.stdout.split("[task ")[1:]
    assert len(blocks) == 3
    for b in blocks[:2]:
        assert "Energy Density            = " in b and "Staggered Susceptibility  = " in b
        assert re.search(r"^Specific Heat: \S+ \+/- \S+$", b, re.M) and re.search(r"^Binder Ratio of Staggered Magnetization: \S+ \+/- \S+$", b, re.M)
        assert "nan" not in b and "inf" not in b
        m = re.search(r"^Number of Clusters: (\S+) \+/- ", b, re.M)
        assert m and abs(float(m.group(1)) - 4.0) < 0.1             # the fake's 3 + step % 3
        assert re.search(r"^Temperature: (0.5|0.25) ", b, re.M)
    assert "Transverse Magnetization" not in blocks[0] and "Transverse Magnetization Density: " in blocks[1]
    # checkpoint: 32 + 256 steps, then resumed into a run three times as long
    ck = os.path.join(str(tmp_path), "run.ck")
    first = subprocess.run([_loop_exe(), "-l", "8", "-t", "0.2", "-n", "256", "--checkpoint", ck], capture_output=True, text=True, env=env)
    assert first.returncode == 0 and open(ck, "rb").read(8) == b"LQCKPT02", first.stderr
    second = subprocess.run([_loop_exe(), "-l", "8", "-t", "0.2", "-n", "768", "--checkpoint", ck], capture_output=True, text=True, env=env)
    assert second.returncode == 0 and "resumed at 0.333333" in second.stdout, (second.stdout, second.stderr)
    third = subprocess.run([_loop_exe(), "-l", "8", "-t", "0.2", "-n", "768", "--checkpoint", ck], capture_output=True, text=True, env=env)
    line = lambda o: re.search(r"Energy Density\s*=\s*(\S+) \+- (\S+)", o.stdout).groups()
    assert third.returncode == 0 and "resumed at 1 " in third.stdout and line(third) == line(second)
    bad = subprocess.run([_loop_exe(), "-l", "10", "-t", "0.2", "-n", "256", "--checkpoint", ck], capture_output=True, text=True, env=env)
    assert bad.returncode != 0 and "lattice size differs" in bad.stderr


def test_loop_driver_two_ranks_rendezvous_with_a_fake_engine(tmp_path):
    """`loop --nranks 2` around the test double: the process forks before it touches the engine, the ranks meet
    through the id file of EACH task, rank 0 alone reports, the id files are gone afterwards."""
    import glob
    env = _fake_engine(tmp_path)
    before = set(glob.glob("/tmp/lq_nccl_id.*"))
    out = subprocess.run([_loop_exe(), "--nranks", "2", "-"], input='L = 4; Jxy = -1; SWEEPS = 128\n{ T = 0.5 } { T = 0.25 }\n',
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.count("[task ") == 2 and out.stdout.count("Energy Density            = ") == 2, out.stdout
    assert set(glob.glob("/tmp/lq_nccl_id.*")) <= before
