"""No GPU needed: the C-ABI library builds, loads and exports every symbol include/vggp.h declares, the ctypes
binding covers all of them, and the product package neither imports the oracle nor has a CPU path."""
import ctypes
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "variational-gridded-gaussian-processes_b200"


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vggp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vggp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    b = importlib.import_module(PKG + ".build")
    path = b.build()
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vggp.h but not exported by libvggp.so"


def test_ctypes_binding_covers_the_header():
    L = importlib.import_module(PKG + "._lib")
    assert sorted(L.SIGNATURES) == header_symbols()
    lib = L.load()
    assert lib.vggp_abi_version() == L.ABI_VERSION
    assert lib.vggp_launch_count() == 0 or lib.vggp_launch_count() > 0


def test_argument_validation_without_a_gpu():
    L = importlib.import_module(PKG + "._lib")
    lib = L.load()
    h = ctypes.c_void_p()
    nk = (ctypes.c_int * 1)(5)
    knots = (ctypes.c_float * 5)(0.0, 0.25, 0.5, 0.75, 1.0)
    ptrs = (ctypes.POINTER(ctypes.c_float) * 1)(ctypes.cast(knots, ctypes.POINTER(ctypes.c_float)))
    assert lib.vggp_plan_create(ctypes.byref(h), 7, 1, nk, ptrs, 0, 0) == -2          # VGGP_E_FAMILY
    assert lib.vggp_plan_create(ctypes.byref(h), 0, 4, nk, ptrs, 0, 0) == -4          # VGGP_E_DIM
    assert lib.vggp_plan_create(ctypes.byref(h), 0, 1, nk, ptrs, 9, 0) == -3          # VGGP_E_DTYPE
    bad = (ctypes.c_float * 5)(0.0, 0.5, 0.25, 0.75, 1.0)
    bptr = (ctypes.POINTER(ctypes.c_float) * 1)(ctypes.cast(bad, ctypes.POINTER(ctypes.c_float)))
    assert lib.vggp_plan_create(ctypes.byref(h), 0, 1, nk, bptr, 0, 0) == -1          # knots not increasing
    assert b"increasing" in lib.vggp_last_error()
    if not torch.cuda.is_available():
        rc = lib.vggp_plan_create(ctypes.byref(h), 0, 1, nk, ptrs, 0, 0)
        assert rc > 0            # a cudaError_t: no device -> the library refuses, it has no CPU path


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, PKG)
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\([\"']oracle", re.M)
    n = 0
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                n += 1
                assert not pat.search(open(os.path.join(base, f)).read()), f"{f} imports the oracle"
    assert n >= 10


def test_product_never_references_the_emulator_and_exports_no_host_path():
    """tests/host_emul (the SIMT emulator) is test infrastructure: nothing in the package, bench.py or __graft_entry__.py
    names it, and the kernels' only concession to it is the VGGP_EMUL guard around the inline-PTX wrappers."""
    pat = re.compile(r"host_emul|emul_lib|fake_cuda|VGGP_HOST_EMUL|libvggp_full_emul")
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for base, _, names in os.walk(os.path.join(ROOT, PKG)):
        files += [os.path.join(base, f) for f in names if f.endswith(".py")]
    for f in files:
        assert not pat.search(open(f).read()), f"{f} references the emulator"
    lib = ctypes.CDLL(importlib.import_module(PKG + ".build").build())
    for sym in ("emul_device_run", "emul_binned_run", "emul_plan_bins"):
        assert not hasattr(lib, sym)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_models_fail_loudly_without_cuda():
    ks = importlib.import_module(PKG + ".models.sparse.kronecker_structure")
    X = torch.rand(10, 2, dtype=torch.float64)
    y = torch.rand(10, dtype=torch.float64)
    model = ks.Matern12B1SplineASVGP(X, y, 5, (0, 1), (0, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        model._elbo()
    basis = importlib.import_module(PKG + ".basis.bspline")
    with pytest.raises(RuntimeError, match="CUDA"):
        basis.B1SplineBasis(torch.linspace(0, 1, 5))(torch.rand(4))
