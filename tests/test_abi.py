"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/arrow_cuda.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from acu import _abi as abi

HEADER = os.path.join(abi.REPO, "include", "arrow_cuda.h")


def declared_symbols(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acu_[a-z0-9_]+)\s*\(", text)) - {"acu_bitmap_bytes"})


def test_generators_live_outside_the_product_library():
    """The synthetic-input generators of the benches / tests are a separate library (include/arrow_cuda_testgen.h)."""
    gen_syms = declared_symbols(os.path.join(abi.REPO, "include", "arrow_cuda_testgen.h"))
    assert gen_syms == sorted(abi.TESTGEN_PROTOTYPES)
    product = C.CDLL(abi.LIB_PATH)
    gen = C.CDLL(abi.TESTGEN_LIB_PATH)
    for s in gen_syms:
        assert not hasattr(product, s) and hasattr(gen, s)


def test_library_is_built():
    assert os.path.exists(abi.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = abi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/arrow_cuda.h but not exported"
        assert s in abi.PROTOTYPES, f"{s} has no ctypes prototype in acu/_abi.py"
    assert set(abi.PROTOTYPES) == set(syms)


def test_struct_layouts_match_header():
    # sizes the C side asserts implicitly (plain pointers and 64-bit integers, no torch types)
    assert C.sizeof(abi.Array) == 56
    assert C.sizeof(abi.ArrayOut) == 40
    assert C.sizeof(abi.ErrorDetail) == 8 + 8 + 8 + 8 + 8 + 256
    assert C.sizeof(abi.Column) == 72 and C.sizeof(abi.ColumnOut) == 64
    # ... and as the library itself was compiled (callable without a GPU)
    lib = abi.load_library()
    for which, t in enumerate([abi.Array, abi.ArrayOut, abi.ErrorDetail, abi.Column, abi.ColumnOut]):
        assert lib.acu_abi_sizeof(which) == C.sizeof(t), t.__name__
    assert lib.acu_abi_sizeof(99) == -1


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA devices ctx creation must fail (never silently compute on the CPU)."""
    lib = abi.load_library()
    h = C.c_void_p()
    st = lib.acu_ctx_create(0, C.byref(h))
    if st == abi.OK:  # a GPU is present: nothing to assert here
        lib.acu_ctx_destroy(h)
        pytest.skip("CUDA device present")
    assert st == abi.ERR_CUDA and not h.value


def _build_c_example(tmp_path):
    """The header is plain C (C11, -pedantic) and the library links from C: examples/hot_path.c."""
    import subprocess
    exe = str(tmp_path / "hot_path")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(abi.REPO, "include"),
           os.path.join(abi.REPO, "examples", "hot_path.c"), "-L" + os.path.dirname(abi.LIB_PATH), "-larrow_cuda", "-larrow_cuda_testgen",
           "-Wl,-rpath," + os.path.dirname(abi.LIB_PATH), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_compiles_and_fails_loudly_without_gpu(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    ctx = C.c_void_p()
    if abi.load_library().acu_ctx_create(0, C.byref(ctx)) == abi.OK:
        abi.load_library().acu_ctx_destroy(ctx)
        pytest.skip("a GPU is present: covered by test_c_example_runs_on_gpu")
    r = subprocess.run([exe, "1000"], capture_output=True, text=True)
    assert r.returncode != 0 and "acu_ctx_create" in r.stderr  # no CPU fallback


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, "3000000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "rows 3000000, selected" in r.stdout
