"""CPU checks of the binned K1 path (csrc/binplan.hpp, csrc/obs_binned.cuh).

tests/host_emul/emul_binned.cpp compiles the SAME __host__ __device__ lane functions the CUDA kernel k_obs_b1_binned
runs (enter cell, per-observation moment update, flush) together with the host planner and the gather index
arithmetic, and runs them sequentially on the CPU.  Here its gradient buffer is compared with a direct float64 numpy
evaluation of the definitions (per-observation mu, p_d, q_d, residual; corner / band scatter), so the run planner, the
padding rules, the moment -> band algebra and the data layout are verified without a GPU.  The emulation is test
infrastructure: the product library has no host path (tests/test_abi_surface.py)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul_binned.cpp")
OUT_DIR = os.path.join(HERE, "host_emul", "_build")
OUT = os.path.join(OUT_DIR, "libvggp_emul.so")
DEPS = [SRC,
        os.path.join(HERE, "..", "variational-gridded-gaussian-processes_b200", "csrc", "obs_binned.cuh"),
        os.path.join(HERE, "..", "variational-gridded-gaussian-processes_b200", "csrc", "binplan.hpp")]


@pytest.fixture(scope="module")
def emul():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-Wno-unknown-pragmas",
                        "-o", OUT, SRC], check=True)
    lib = C.CDLL(OUT)
    lib.emul_binned_run.restype = C.c_int
    lib.emul_plan_bins.restype = C.c_int
    return lib


def make_problem(D, knots_per_dim, n, dtype, seed, frac_outside=0.05, clustered=True, on_knots=True):
    rng = np.random.default_rng(seed)
    meshes = []
    for d in range(D):
        K = knots_per_dim[d]
        if d == 0:
            t = np.linspace(0.0, 1.0, K, dtype=np.float32)
        else:                                           # non-uniform mesh
            t = np.cumsum(rng.uniform(0.5, 1.5, K)).astype(np.float32)
            t = ((t - t[0]) / (t[-1] - t[0]) * (1.0 + d)).astype(np.float32)
        meshes.append(t)
    X = np.empty((n, D), dtype=dtype)
    for d in range(D):
        lo, hi = float(meshes[d][0]), float(meshes[d][-1])
        span = hi - lo
        if clustered:            # very uneven cell populations: most points in a few cells
            centers = rng.uniform(lo, hi, 3)
            which = rng.integers(0, 4, n)
            x = np.where(which < 3, centers[np.minimum(which, 2)] + 0.03 * span * rng.standard_normal(n),
                         rng.uniform(lo, hi, n))
        else:
            x = rng.uniform(lo, hi, n)
        out = rng.random(n) < frac_outside
        x = np.where(out, np.where(rng.random(n) < 0.5, lo - 0.1 * span * rng.random(n) - 1e-3,
                                   hi + 0.1 * span * rng.random(n) + 1e-3), np.clip(x, lo, hi))
        X[:, d] = x.astype(dtype)
    if on_knots and n >= 8:      # exact knot hits: first, last and interior knots
        for d in range(D):
            X[0, d] = meshes[d][0]
            X[1, d] = meshes[d][-1]
            X[2, d] = meshes[d][knots_per_dim[d] // 2]
            X[3, d] = meshes[d][1]
    y = (np.sin(3 * X.astype(np.float64).sum(axis=1)) + 0.1 * rng.standard_normal(n)).astype(dtype)
    M = int(np.prod(knots_per_dim))
    alpha = (0.5 * rng.standard_normal(M)).astype(dtype)
    bands = []
    for d in range(D):
        K = knots_per_dim[d]
        bands.append(dict(pd=rng.uniform(0.5, 2.0, K), po=rng.uniform(-0.5, 0.5, K - 1),
                          qd=rng.uniform(0.1, 1.0, K), qo=rng.uniform(-0.2, 0.2, K - 1)))
    return meshes, X, y, alpha, bands


def cell_tables(meshes, bands, dtype):
    """[pe0 pe1 pe2 qe0 qe1 qe2 h rh] x K per dimension, as k_fwd_reduce / k_fwd_qtable write them (grid.cuh)."""
    tabs, tab_off, off = [], [], 0
    for t, b in zip(meshes, bands):
        K = t.size
        tab = np.zeros((8, K), dtype=np.float64)
        for name, r0 in (("p", 0), ("q", 3)):
            A = b[name + "d"].copy()
            B2 = np.zeros(K)
            B2[:-1] = 2.0 * b[name + "o"]
            Cn = np.zeros(K)
            Cn[:-1] = b[name + "d"][1:]
            tab[r0] = A
            tab[r0 + 1] = B2 - 2.0 * A
            tab[r0 + 2] = A - B2 + Cn
            tab[r0 + 1, -1] = -2.0 * A[-1]          # last entry: B2 = C = 0 (never addressed by a cell)
            tab[r0 + 2, -1] = A[-1]
        h32 = np.ones(K, dtype=np.float32)
        h32[:-1] = t[1:] - t[:-1]                   # float32 knot difference, reference semantics
        h = h32.astype(dtype)
        tab[6] = h
        tab[7] = (np.ones(K, dtype=dtype) / h).astype(dtype)
        tabs.append(tab.astype(dtype).reshape(-1))
        tab_off.append(off)
        off += 8 * K
    return np.concatenate(tabs), tab_off


def expected(meshes, X, y, alpha, bands, dtype):
    """Direct float64 evaluation of the definitions (DESIGN.md appendix A / obs.cuh header)."""
    n, D = X.shape
    Ks = [t.size for t in meshes]
    strides = [int(np.prod(Ks[d + 1:])) for d in range(D)]
    Xd = X.astype(np.float64)
    cs, ws, inside = [], [], np.ones(n, dtype=bool)
    for d in range(D):
        t = meshes[d].astype(dtype)
        c = np.clip(np.searchsorted(t, X[:, d], side="left") - 1, 0, Ks[d] - 2)
        inside &= (X[:, d] >= t[0]) & (X[:, d] <= t[-1])
        h = (meshes[d][1:] - meshes[d][:-1]).astype(np.float64)            # float32 difference, then promoted
        a = (Xd[:, d] - meshes[d].astype(np.float64)[c]) / h[c]
        cs.append(c)
        ws.append(a)
    idx = np.nonzero(inside)[0]
    galpha = np.zeros(alpha.size)
    gband = [np.zeros((4, K)) for K in Ks]
    al = alpha.astype(np.float64)
    mu = np.zeros(n)
    for corner in range(1 << D):
        wt = np.ones(n)
        off = np.zeros(n, dtype=np.int64)
        for d in range(D):
            hi = (corner >> (D - 1 - d)) & 1
            wt *= ws[d] if hi else (1.0 - ws[d])
            off += (cs[d] + hi) * strides[d]
        mu += wt * al[off]
    r = y.astype(np.float64) - mu
    p, q = [], []
    for d in range(D):
        b, c, a = bands[d], cs[d], ws[d]
        p.append((1 - a) ** 2 * b["pd"][c] + 2 * a * (1 - a) * b["po"][c] + a ** 2 * b["pd"][c + 1])
        q.append((1 - a) ** 2 * b["qd"][c] + 2 * a * (1 - a) * b["qo"][c] + a ** 2 * b["qd"][c + 1])
    for corner in range(1 << D):
        wt = np.ones(n)
        off = np.zeros(n, dtype=np.int64)
        for d in range(D):
            hi = (corner >> (D - 1 - d)) & 1
            wt *= ws[d] if hi else (1.0 - ws[d])
            off += (cs[d] + hi) * strides[d]
        np.add.at(galpha, off[idx], (r * wt)[idx])
    for d in range(D):
        op = np.ones(n)
        oq = np.ones(n)
        for e in range(D):
            if e != d:
                op *= p[e]
                oq *= q[e]
        a, c = ws[d], cs[d]
        for row, o in ((0, op), (2, oq)):
            np.add.at(gband[d][row], c[idx], (o * (1 - a) ** 2)[idx])
            np.add.at(gband[d][row], c[idx] + 1, (o * a ** 2)[idx])
            np.add.at(gband[d][row + 1], c[idx], (o * a * (1 - a))[idx])
    pp = np.prod(np.stack(p), axis=0)
    qq = np.prod(np.stack(q), axis=0)
    E = float(np.sum((r ** 2 - pp + qq)[idx]) + np.sum(y.astype(np.float64)[~inside] ** 2))
    return galpha, np.concatenate([g.reshape(-1) for g in gband]), E, int(inside.sum())


def run_emul(lib, meshes, X, y, alpha, bands, dtype, run_cap):
    n, D = X.shape
    Ks = [int(t.size) for t in meshes]
    strides = [int(np.prod(Ks[d + 1:])) for d in range(D)]
    tab, tab_off = cell_tables(meshes, bands, dtype)
    knots = np.concatenate(meshes).astype(np.float32)
    knot_off = [int(sum(Ks[:d])) for d in range(D)]
    band_off = [int(4 * sum(Ks[:d])) for d in range(D)]
    xs = [np.ascontiguousarray(X[:, d]) for d in range(D)]
    xptr = (C.c_void_p * D)(*[x.ctypes.data for x in xs])
    galpha = np.zeros(alpha.size, dtype=dtype)
    gband = np.zeros(4 * sum(Ks), dtype=dtype)
    gs = np.zeros(2, dtype=np.float64)
    stats = np.zeros(6, dtype=np.int64)
    iarr = lambda v: (C.c_int * D)(*v)
    rc = lib.emul_binned_run(C.c_int(0 if dtype == np.float32 else 1), C.c_int(D), iarr(Ks),
                             knots.ctypes.data_as(C.POINTER(C.c_float)), xptr, C.c_void_p(y.ctypes.data),
                             C.c_int64(n), C.c_int(run_cap), iarr(strides), iarr(band_off), iarr(tab_off),
                             iarr(knot_off), C.c_void_p(tab.ctypes.data), C.c_void_p(alpha.ctypes.data),
                             C.c_void_p(galpha.ctypes.data), C.c_void_p(gband.ctypes.data),
                             gs.ctypes.data_as(C.POINTER(C.c_double)), stats.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    return galpha, gband, gs, stats


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))


CASES = [
    # D, knots, n, run_cap
    (1, (9,), 700, 64),
    (1, (33,), 5000, 16),
    (2, (7, 5), 3000, 32),
    (2, (17, 12), 20000, 256),
    (2, (6, 9), 4000, 4),
    (3, (5, 4, 6), 6000, 48),
]


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 3e-4)])
@pytest.mark.parametrize("D,knots,n,run_cap", CASES)
def test_binned_lane_code_matches_definitions(emul, D, knots, n, run_cap, dtype, tol):
    meshes, X, y, alpha, bands = make_problem(D, knots, n, dtype, seed=D * 1000 + n + run_cap)
    ga, gb, gs, stats = run_emul(emul, meshes, X, y, alpha, bands, dtype, run_cap)
    ea, eb, eE, n_in = expected(meshes, X, y, alpha, bands, dtype)
    assert stats[0] == n_in                         # cell membership (bit-exact rule) agrees with searchsorted
    assert gs[1] == n
    assert rel(ga, ea) < tol
    assert rel(gb, eb) < tol
    assert abs(gs[0] - eE) / abs(eE) < tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_binned_edge_cases(emul, dtype):
    # no observations; all outside; a single observation; a single full cell larger than the cap
    meshes, X, y, alpha, bands = make_problem(2, (6, 5), 64, dtype, seed=5, on_knots=False)
    tol = 1e-11 if dtype == np.float64 else 3e-4
    ga, gb, gs, st = run_emul(emul, meshes, X[:0], y[:0], alpha, bands, dtype, 32)
    assert st[2] == 0 and not ga.any() and not gb.any() and gs[0] == 0.0 and gs[1] == 0
    Xo = X.copy()
    Xo[:, 0] = 7.0
    ga, gb, gs, st = run_emul(emul, meshes, Xo, y, alpha, bands, dtype, 32)
    assert st[0] == 0 and st[2] == 0 and not ga.any() and not gb.any()
    assert abs(gs[0] - float(np.sum(y.astype(np.float64) ** 2))) < 1e-5 * max(1.0, gs[0])
    for sub in (slice(0, 1), slice(0, 64)):
        Xs, ys = X[sub].copy(), y[sub].copy()
        if sub.stop == 64:                           # everything into one cell
            Xs[:, 0] = meshes[0][2] + 0.5 * (meshes[0][3] - meshes[0][2]) * np.linspace(0.01, 0.99, 64).astype(dtype)
            Xs[:, 1] = meshes[1][1] + 0.5 * (meshes[1][2] - meshes[1][1]) * np.linspace(0.99, 0.01, 64).astype(dtype)
        ga, gb, gs, st = run_emul(emul, meshes, Xs, ys, alpha, bands, dtype, 12)
        ea, eb, eE, n_in = expected(meshes, Xs, ys, alpha, bands, dtype)
        assert st[0] == n_in
        assert rel(ga, ea) < tol and rel(gb, eb) < tol and abs(gs[0] - eE) <= tol * max(abs(eE), 1.0)


def test_plan_bins_properties(emul):
    rng = np.random.default_rng(0)
    for D, cap in ((1, 4), (2, 64), (3, 256), (2, 255)):
        ncells = 1000
        count = np.zeros(ncells + 1, dtype=np.uint32)
        count[:ncells] = rng.poisson(3, ncells) * (rng.random(ncells) < 0.6)
        count[7] = 5000                                # one very crowded cell
        count[ncells] = 123                            # outside
        out = np.zeros(8, dtype=np.int64)
        run_n = np.zeros(32 * 4096, dtype=np.int32)
        rc = emul.emul_plan_bins(count.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(ncells), C.c_int(cap), C.c_int(D),
                                 out.ctypes.data_as(C.POINTER(C.c_int64)), run_n.ctypes.data_as(C.POINTER(C.c_int32)),
                                 C.c_int64(run_n.size))
        assert rc == 0
        n, n_in, n_runs, n_tasks, data_elems, rmax, rmin, tot = out
        cap4 = cap // 4 * 4
        assert n == count.sum() and n_in == count[:ncells].sum() and tot == n_in
        assert n_tasks == (n_runs + 31) // 32
        lens = run_n[: 32 * n_tasks]
        assert (lens[:n_runs] > 0).all() and (lens[n_runs:] == 0).all()
        assert lens.max() <= cap4 and rmax <= cap4 and rmax % 4 == 0 and rmin % 4 == 0
        assert (np.diff(lens[:n_runs]) <= 0).all()                      # longest first
        expect_runs = sum(-(-int(c) // cap4) for c in count[:ncells] if c > 0)
        assert n_runs == expect_runs
        per_task = lens.reshape(n_tasks, 32)
        R = (per_task[:, 0] + 3) // 4 * 4
        assert data_elems == int((32 * R * (D + 1)).sum())
    bad = np.zeros(2, dtype=np.uint32)
    out = np.zeros(8, dtype=np.int64)
    assert emul.emul_plan_bins(bad.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(1), C.c_int(3), C.c_int(2),
                               out.ctypes.data_as(C.POINTER(C.c_int64)), None, C.c_int64(0)) != 0
