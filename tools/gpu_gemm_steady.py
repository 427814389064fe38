"""Steady-state rate of k_gemm_group / k_gemm_one (float64 DMMA) against the measured DMMA peak, next to the latency-bound
511^3 mode product of configs[2]:  python tools/gpu_gemm_steady.py  -> one JSON line per shape."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vggp_b200 as vg  # noqa: E402


def main():
    lib = vg._lib.load()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_dmma_peak.json")))["dmma_f64_tflops"]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    shapes = [(1, 511, 511, 511, 1), (16, 511, 511, 511, 1), (1, 2048, 2048, 2048, 1), (1, 4096, 4096, 1024, 1), (64, 512, 512, 512, 1)]
    if len(sys.argv) > 1 and sys.argv[1] == "--big-only":
        shapes = [(1, 4096, 4096, 1024, 1)]
    for batch, m, n, k, sk in shapes:
        A = torch.randn(batch, m, k, dtype=torch.float64, device=dev)
        B = torch.randn(batch, k, n, dtype=torch.float64, device=dev)
        C = torch.zeros(batch, m, n, dtype=torch.float64, device=dev)

        def run():
            rc = lib.vggp_gemm_f64(1, batch, m, n, k, ctypes.c_double(1.0), ctypes.c_void_p(A.data_ptr()), k, 1, m * k if batch > 1 else 0,
                                   ctypes.c_void_p(B.data_ptr()), n, 1, k * n if batch > 1 else 0, ctypes.c_double(0.0),
                                   ctypes.c_void_p(C.data_ptr()), n, 1, m * n, sk, st)
            assert rc == 0, rc
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        ref = torch.matmul(A[0], B[0])
        err = ((C[0] - ref).norm() / ref.norm()).item()
        tf = 2.0 * batch * m * n * k / (ms * 1e-3) / 1e12
        print(json.dumps({"batch": batch, "m": m, "n": n, "k": k, "splitk_arg": sk, "ms": ms, "tflops": tf, "frac_of_dmma_peak": tf / peak,
                          "rel_err_vs_torch": err}), flush=True)


if __name__ == "__main__":
    main()
