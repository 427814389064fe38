"""Randomised campaigns under the SIMT emulator (tests/host_emul), CPU only.  Longer-running than the test suite; this is
how the first-knot bug of k_obs_b1 was found.

    python tools/emul_stress.py device --seconds 120         # K1 kernels of all four layouts vs the float64 definitions
    python tools/emul_stress.py step   --seconds 240         # whole B1 steps through the C ABI vs the oracle
    python tools/emul_stress.py b0     --seconds 200         # whole B0 steps (dense-feature kernel and scan form) vs the oracle
    python tools/emul_stress.py device --asan --seconds 240  # the same with AddressSanitizer (re-executes itself with libasan preloaded)
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "host_emul")


def device_campaign(seconds, lib_path, seed):
    import test_device_emul as T
    from test_binned_host_emul import expected, make_problem, rel
    lib = C.CDLL(lib_path)
    lib.emul_device_run.restype = C.c_int
    rng = np.random.default_rng(seed)
    t0, k = time.time(), 0
    while time.time() - t0 < seconds:
        D = int(rng.integers(1, 4))
        knots = tuple(int(rng.integers(3, 14)) for _ in range(D))
        n = int(rng.integers(1, 9000))
        cap = int(rng.choice([4, 8, 12, 32, 100, 256, 1000]))
        dtype = [np.float64, np.float32][int(rng.integers(0, 2))]
        layout = ["binned_ldg", "binned_tma", "packed_sorted", "packed_unsorted"][int(rng.integers(0, 4))]
        args = dict(seed=int(rng.integers(0, 1 << 30)), frac_outside=float(rng.choice([0, 0.05, 0.5])),
                    clustered=bool(rng.integers(0, 2)), on_knots=n >= 8)
        meshes, X, y, alpha, bands = make_problem(D, knots, n, dtype, **args)
        ga, gb, gs, st = T.run_device(lib, layout, meshes, X, y, alpha, bands, dtype, cap, blocks_cap=int(rng.integers(1, 5)))
        ea, eb, eE, n_in = expected(meshes, X, y, alpha, bands, dtype)
        tol = 1e-10 if dtype == np.float64 else 5e-4
        ra = rel(ga, ea) if np.linalg.norm(ea) > 0 else float(np.abs(ga).max())
        rb = rel(gb, eb) if np.linalg.norm(eb) > 0 else float(np.abs(gb).max())
        re_ = abs(gs[0] - eE) / max(abs(eE), 1.0)
        if not ((layout.startswith("packed") or st[2] == n_in) and gs[1] == n and ra < tol and rb < tol and re_ < tol):
            print("FAIL", dict(D=D, knots=knots, n=n, cap=cap, dtype=dtype.__name__, layout=layout, **args), ra, rb, re_)
            return 1
        k += 1
    print("device campaign:", k, "configurations, no discrepancy")
    return 0


def step_campaign(seconds, family, seed):
    import torch
    import emul_lib
    from oracle import vggp_oracle as O
    from test_gpu_elbo import make_problem, oracle_value_and_grads
    lib, L = emul_lib.load()
    ofam, fam = (O.B1_ASVGP, L.B1_ASVGP) if family == "B1" else (O.B0_GRIDDED, L.B0_GRIDDED)
    layouts = ["raw", "packed_sorted", "packed_unsorted", "binned"] if family == "B1" else ["raw", "binned"]
    rng = np.random.default_rng(seed)

    def rel(a, b):
        a = np.asarray(a, dtype=np.float64).ravel()
        b = b.detach().double().numpy().ravel()
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    t0, k = time.time(), 0
    while time.time() - t0 < seconds:
        D = int(rng.integers(1, 4 if family == "B1" else 3))
        knots = tuple(int(rng.integers(3, 20)) for _ in range(D))
        if np.prod(knots) > 1500:
            continue
        N = int(rng.integers(1, 1500))
        layout = layouts[int(rng.integers(0, len(layouts)))]
        seed_p = int(rng.integers(0, 1 << 30))
        lo, hi = [(-0.05, 1.05), (0.0, 1.0), (-1.0, 2.0)][int(rng.integers(0, 3))]
        meshes, X, y, l, s2, noise, m, Ls = make_problem(knots, N, seed=seed_p, family=ofam, x_lo=lo, x_hi=hi)
        scale = float(rng.choice([1.0, 2.5]))
        try:
            elbo_ref, g_ref = oracle_value_and_grads(ofam, meshes, X, y, l, s2, noise, m, Ls, scale=scale)
        except Exception:
            continue
        plan = emul_lib.EmuPlan(lib, L, fam, [t.numpy() for t in meshes], np.float64)
        theta = torch.cat([l, s2, noise.reshape(1)]).numpy().copy()
        mm, Lc = m.numpy().copy(), torch.cat([Lx.reshape(-1) for Lx in Ls]).numpy().copy()
        xs, yy = [np.ascontiguousarray(X[:, d].numpy()) for d in range(D)], y.numpy().copy()
        if layout == "raw":
            out = plan.step(theta, mm, Lc, xs, yy, scale)
        elif layout == "binned":
            out = plan.step(theta, mm, Lc, plan.bin(xs, yy, int(rng.choice([4, 32, 256]))), None, scale)
        else:
            out = plan.step(theta, mm, Lc, plan.pack(xs, yy, layout == "packed_sorted"), None, scale)
        info = plan.read_info()
        m_per_dim = plan.m_per_dim
        plan.close()
        if info != 0:
            continue
        errs = [abs(out[0][0] - elbo_ref.item()) / abs(elbo_ref.item()), rel(out[1][:D], g_ref[0]), rel(out[1][D:2 * D], g_ref[1]),
                rel(out[1][2 * D], g_ref[2]), rel(out[2], g_ref[3])]
        off = 0
        for d, n in enumerate(m_per_dim):
            errs.append(rel(np.tril(out[3][off:off + n * n].reshape(n, n)), torch.tril(g_ref[4 + d])))
            off += n * n
        if max(errs) > 1e-6 or out[0][3] != N:
            print("FAIL", dict(family=family, knots=knots, N=N, layout=layout, seed=seed_p, lo=lo, hi=hi, scale=scale), ["%.1e" % e for e in errs])
            return 1
        k += 1
    print(f"{family} step campaign:", k, "configurations, no discrepancy")
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["device", "step", "b0"])
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--asan", action="store_true", help="device campaign only: AddressSanitizer build of the harness")
    args = ap.parse_args()
    if args.what == "device":
        import test_device_emul as T
        lib_path = T.OUT
        if args.asan:
            lib_path = os.path.join(EMU, "_build", "libvggp_device_emul_asan.so")
            if "libasan" not in os.environ.get("LD_PRELOAD", ""):
                os.makedirs(os.path.dirname(lib_path), exist_ok=True)
                subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-extern-tls-init", "-shared",
                                "-fsanitize=address", "-fno-omit-frame-pointer", "-Wno-unknown-pragmas", "-Wno-attributes",
                                "-I", os.path.join(EMU, "fake_cuda"), "-o", lib_path] + T.SRCS, check=True)
                asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
                env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")
                sys.exit(subprocess.run([sys.executable] + sys.argv, env=env).returncode)
        elif not os.path.exists(lib_path):
            subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_device_emul.py"), "-q", "-k", "edge"], check=True)
        sys.exit(device_campaign(args.seconds, lib_path, args.seed))
    sys.exit(step_campaign(args.seconds, "B1" if args.what == "step" else "B0", args.seed))


if __name__ == "__main__":
    main()
