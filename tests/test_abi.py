"""CPU tests of the drop-in boundary: liblq.so loads without a GPU, exports exactly the symbols
include/lq.h declares, the ctypes mirrors have the C layout, and computing entry points fail
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lq.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import looper_b200 as lq
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lq.lib, n), f"{n} declared in include/lq.h but not exported by liblq.so"
    assert sorted(lq.EXPORTS) == names
    nm = subprocess.run(["nm", "-D", "--defined-only", lq.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lq_[a-z0-9_]+)", nm))
    assert set(names) <= exported


def test_struct_layouts_match_header():
    import looper_b200 as lq
    assert C.sizeof(lq.LqOp) == 16 and lq.OP_DTYPE.itemsize == 16
    assert C.sizeof(lq.LqCollector) == 20 * 8 and lq.COLLECTOR_DTYPE.itemsize == 20 * 8
    assert C.sizeof(lq.LqTimer) == 56
    assert C.sizeof(lq.LqModel) == 8 + 32 + 8 + 8 + 8
    assert C.sizeof(lq.LqLattice) == 8 + 3 * 8 + 16 + 8
    assert C.sizeof(lq.LqOptions) == 8 + 4 + 4 + 8 * 3 + 4 * 5 + 4   # seed, device, tile_sites, 3 doubles, rank..cut, tail padding
    # compile the header as C and compare sizeof with the compiler's view
    prog = r'''
    #include <stdio.h>
    #include "lq.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(lq_op), sizeof(lq_collector),
      sizeof(lq_timer), sizeof(lq_model), sizeof(lq_lattice), sizeof(lq_options), sizeof(lq_info)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o",
                               os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split()
    got = [C.sizeof(x) for x in (lq.LqOp, lq.LqCollector, lq.LqTimer, lq.LqModel, lq.LqLattice,
                                 lq.LqOptions, lq.LqInfo)]
    assert [int(x) for x in out] == got


def test_no_cpu_fallback_without_a_device():
    import looper_b200 as lq
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the negative test is for the CPU box")
    with pytest.raises(lq.LqError) as e:
        lq.Engine(lq.chain_lattice(8), 5.0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_argument_validation_needs_no_device():
    import looper_b200 as lq
    h = C.c_void_p()
    assert lq.lib.lq_create(C.byref(h), None, None, 1.0, None) == -1
    assert lq.lib.lq_sweep(None, None) == -1
    assert lq.lib.lq_kernel_launches(None) == 0
    assert b"sm_100a" in lq.lib.lq_version()


def test_product_path_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under alps-looper_b200/ or include/ may name it"""
    bad = []
    for base in ("alps-looper_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle|liboracle|orc_", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_lattice_generators_and_weights():
    import looper_b200 as lq
    lat = lq.hypercubic_lattice((4, 6))
    assert lat["num_sites"] == 24 and len(lat["src"]) == 48
    deg = np.bincount(np.concatenate([lat["src"], lat["dst"]]), minlength=24)
    assert np.all(deg == 4)
    g = lat["gauge"]
    assert np.all(g[lat["src"]] * g[lat["dst"]] == -1)          # bipartite
    lat3 = lq.hypercubic_lattice((4, 4, 4))
    assert len(lat3["src"]) == 3 * 64
    lad = lq.hypercubic_lattice((8, 2))
    assert len(lad["src"]) == 16 + 8                            # ring of length 2 has one rung bond
    ch = lq.chain_lattice(8)
    assert list(ch["dst"]) == [1, 2, 3, 4, 5, 6, 7, 0]          # standalone/common.h:92-93
    v, off, sign = lq.xxz_weights(1.0, 1.0)
    assert v == [0.5, 0, 0, 0] and off == 0.25 and sign == -1    # test/weight.op:9
    v, off, sign = lq.xxz_weights(1.0, 0.5)
    assert v == [0.375, 0.125, 0, 0] and off == 0.25             # test/weight.op:12
    v, off, sign = lq.xxz_weights(1.0, 1.0, 0.1)
    assert np.allclose(v, [0.45, 0.05, 0.1, 0]) and off == pytest.approx(0.3)   # test/weight.op:10
