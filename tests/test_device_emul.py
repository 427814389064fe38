"""The library's CUDA kernels executed on the CPU under a SIMT emulator (tests/host_emul/fake_cuda + emul_core.cpp).

g++ compiles csrc/obs.cuh and csrc/obs_binned.cuh unchanged against a stand-in <cuda_runtime.h>: the threads of a
block are cooperatively scheduled fibres, `__shared__` is block-wide storage, warp shuffles / __syncthreads / atomics /
mbarrier waits are scheduling points, and bulk async copies (TMA) complete late and out of order.  The emulator is
validated on k_obs_b1, the kernel that passed the GPU parity suite on B200 (layouts 0 and 1 below); the same harness
then runs the kernels that have NOT been on a GPU yet -- k_bin_histogram, k_bin_gather, k_bin_sum_y2,
k_obs_b1_binned (LDG stream) and k_obs_b1_binned_tma (TMA ring) -- and compares their gradient buffer with the float64
numpy evaluation of the definitions from tests/test_binned_host_emul.py.  Test infrastructure only: the product has no
CPU path; this checks indexing, protocols and arithmetic of device code, not memory ordering or speed."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from test_binned_host_emul import cell_tables, expected, make_problem, rel

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emul")
OUT_DIR = os.path.join(EMU, "_build")
OUT = os.path.join(OUT_DIR, "libvggp_device_emul.so")
CSRC = os.path.join(HERE, "..", "variational-gridded-gaussian-processes_b200", "csrc")
SRCS = [os.path.join(EMU, "emul_core.cpp"), os.path.join(EMU, "emul_device.cpp")]
DEPS = SRCS + [os.path.join(EMU, "fake_cuda", "cuda_runtime.h")] + [
    os.path.join(CSRC, f) for f in ("obs.cuh", "obs_binned.cuh", "binplan.hpp", "common.cuh")]

LAYOUTS = {"packed_sorted": 0, "packed_unsorted": 1, "binned_ldg": 2, "binned_tma": 3}


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        # -fno-extern-tls-init: `extern __shared__` becomes a block-scope `extern thread_local`; without the flag gcc
        # routes it through a weak TLS-init wrapper that is null here
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-extern-tls-init", "-shared",
                        "-Wno-unknown-pragmas", "-Wno-attributes", "-I", os.path.join(EMU, "fake_cuda"), "-o", OUT] + SRCS, check=True)
    lib = C.CDLL(OUT)
    lib.emul_device_run.restype = C.c_int
    return lib


def run_device(lib, layout, meshes, X, y, alpha, bands, dtype, run_len, blocks_cap=3):
    n, D = X.shape
    Ks = [int(t.size) for t in meshes]
    M = int(np.prod(Ks))
    tab, _ = cell_tables(meshes, bands, dtype)
    knots = np.concatenate(meshes).astype(np.float32)
    xs = [np.ascontiguousarray(X[:, d]) for d in range(D)]
    xptr = (C.c_void_p * D)(*[x.ctypes.data for x in xs])
    esz = np.dtype(dtype).itemsize
    n_elems = M + 4 * sum(Ks)
    soff = (n_elems * esz + 7) // 8 * 8
    gbuf = np.zeros(soff + 64, dtype=np.uint8)
    stats = np.zeros(8, dtype=np.int64)
    rc = lib.emul_device_run(C.c_int(0 if dtype == np.float32 else 1), C.c_int(D), C.c_int(LAYOUTS[layout]),
                             (C.c_int * D)(*Ks), knots.ctypes.data_as(C.POINTER(C.c_float)), xptr,
                             C.c_void_p(y.ctypes.data), C.c_int64(n), C.c_int(run_len), C.c_int(blocks_cap),
                             C.c_void_p(tab.ctypes.data), C.c_void_p(alpha.ctypes.data), C.c_void_p(gbuf.ctypes.data),
                             C.c_int64(soff), stats.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    obs = gbuf[: n_elems * esz].view(dtype)
    gs = gbuf[soff: soff + 64].view(np.float64)
    return obs[:M].copy(), obs[M:].copy(), gs.copy(), stats


CASES = [
    # D, knots, n, run length (packed) / run cap (binned)
    (1, (9,), 900, 16),
    (2, (7, 5), 2500, 32),
    (2, (12, 9), 6000, 64),
    (3, (5, 4, 6), 3000, 24),
]


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 3e-4)])
@pytest.mark.parametrize("D,knots,n,run_len", CASES)
def test_device_kernels_under_the_emulator(emu, D, knots, n, run_len, dtype, tol, layout):
    meshes, X, y, alpha, bands = make_problem(D, knots, n, dtype, seed=7 * D + n)
    ga, gb, gs, stats = run_device(emu, layout, meshes, X, y, alpha, bands, dtype, run_len)
    ea, eb, eE, n_in = expected(meshes, X, y, alpha, bands, dtype)
    assert gs[1] == n
    assert rel(ga, ea) < tol
    assert rel(gb, eb) < tol
    assert abs(gs[0] - eE) <= tol * max(abs(eE), 1.0)
    if layout.startswith("binned"):
        assert stats[2] == n_in


@pytest.mark.parametrize("layout", ["binned_ldg", "binned_tma"])
def test_binned_device_edge_cases(emu, layout):
    dtype = np.float64
    meshes, X, y, alpha, bands = make_problem(2, (6, 5), 200, dtype, seed=11, on_knots=False)
    # all outside: no task, one block adds sum y^2 and n
    Xo = X.copy()
    Xo[:, 1] = -5.0
    ga, gb, gs, st = run_device(emu, layout, meshes, Xo, y, alpha, bands, dtype, 32)
    assert st[0] == 0 and not ga.any() and not gb.any() and gs[1] == 200
    assert abs(gs[0] - float(np.sum(y ** 2))) < 1e-9 * gs[0]
    # one observation; one crowded cell split into many runs (odd group counts exercise the partial TMA stage)
    for Xs, ys, cap in ((X[:1], y[:1], 32), (None, None, 12), (None, None, 20)):
        if Xs is None:
            Xs, ys = X.copy(), y.copy()
            Xs[:, 0] = meshes[0][2] + 0.5 * (meshes[0][3] - meshes[0][2]) * np.linspace(0.01, 0.99, 200)
            Xs[:, 1] = meshes[1][1] + 0.5 * (meshes[1][2] - meshes[1][1]) * np.linspace(0.99, 0.01, 200)
        ga, gb, gs, st = run_device(emu, layout, meshes, Xs, ys, alpha, bands, dtype, cap, blocks_cap=2)
        ea, eb, eE, n_in = expected(meshes, Xs, ys, alpha, bands, dtype)
        assert st[2] == n_in
        assert rel(ga, ea) < 1e-11 and rel(gb, eb) < 1e-11 and abs(gs[0] - eE) <= 1e-11 * max(abs(eE), 1.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", ["packed_unsorted", "packed_sorted", "binned_ldg"])
def test_first_knot_hit_followed_by_a_point_left_of_the_mesh(emu, layout, dtype):
    """Regression (found by randomised emulator runs): in an unsorted stream, k_obs_b1 used to widen the cached lower
    bound of cell 0 to -inf after an observation exactly on the first knot, so an observation LEFT of the mesh that
    followed in the same lane run was taken as inside cell 0.  The bound is now lowered by one ulp only."""
    rng = np.random.default_rng(5)
    meshes, X, y, alpha, bands = make_problem(1, (5,), 64, dtype, seed=9, frac_outside=0.0, clustered=False, on_knots=False)
    t0, t1 = float(meshes[0][0]), float(meshes[0][1])
    X[:, 0] = (t0 + (t1 - t0) * rng.uniform(0.1, 0.9, 64)).astype(dtype)      # everything in cell 0 ...
    X[10, 0] = dtype(t0)                                                        # ... one exact first-knot hit ...
    X[11:20, 0] = dtype(t0) - dtype(0.05) * np.arange(1, 10, dtype=dtype)       # ... then points left of the mesh
    ga, gb, gs, _ = run_device(emu, layout, meshes, X, y, alpha, bands, dtype, 32, blocks_cap=1)
    ea, eb, eE, n_in = expected(meshes, X, y, alpha, bands, dtype)
    assert n_in == 64 - 9
    tol = 1e-11 if dtype == np.float64 else 3e-4
    assert rel(ga, ea) < tol and rel(gb, eb) < tol and abs(gs[0] - eE) <= tol * abs(eE)
