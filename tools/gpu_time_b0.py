"""GPU timing of the B0 (cell-integrated) family: dense-feature kernel k_obs_b0 against the scan form k_obs_b0s
(DESIGN.md section 10).  Run on a B200 after the gated tests pass:
    python tools/gpu_time_b0.py [--knots 129] [--n 1048576]
Prints one JSON line per variant: ms per ELBO value+gradient step and the relative difference of the results."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vggp_b200 as vg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--knots", type=int, default=129)
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--skip-dense", action="store_true", help="the dense kernel is O(N M): skip it at large sizes")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    meshes = [torch.linspace(0, 1, args.knots)] * 2
    X = (torch.rand(args.n, 2, generator=g, dtype=torch.float64) * 1.1 - 0.05).to(torch.float32)
    y = (torch.sin(5 * X[:, 0]) + torch.cos(7 * X[:, 1]) + 0.05 * torch.randn(args.n, generator=g)).to(torch.float32)
    M1 = args.knots - 1
    plan = vg.GridPlan(vg.B0_GRIDDED, meshes, torch.float32, dev)
    theta = torch.tensor([0.08, 0.08, 1.2, 1.2, 0.01], dtype=torch.float64, device=dev)
    m = (0.1 * torch.randn(M1 * M1, generator=g, dtype=torch.float64)).to(dev)
    L = torch.cat([(0.5 * torch.eye(M1, dtype=torch.float64)).reshape(-1)] * 2).to(dev)
    xs = [X[:, d].contiguous().to(dev) for d in range(2)]
    yd = y.to(dev)
    results = {}
    variants = [("scan", plan.bin(xs, yd, run_cap=256), None)]
    if not args.skip_dense:
        variants.append(("dense", xs, yd))
    for name, obs, yy in variants:
        for _ in range(2):
            out = plan.step(theta, m, L, obs, yy)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            out = plan.step(theta, m, L, obs, yy)
        b.record()
        torch.cuda.synchronize()
        results[name] = [t.clone() for t in out]
        print(json.dumps({"variant": name, "knots": args.knots, "n": args.n, "ms_per_step": a.elapsed_time(b) / args.steps,
                          "elbo": out[0][0].item(), "info": plan.read_info()}), flush=True)
    if "dense" in results:
        rel = lambda u, v: ((u - v).norm() / v.norm()).item()
        print(json.dumps({"elbo_rel_diff": abs(results["scan"][0][0].item() - results["dense"][0][0].item())
                          / abs(results["dense"][0][0].item()),
                          "dtheta_rel_diff": rel(results["scan"][1], results["dense"][1]),
                          "dm_rel_diff": rel(results["scan"][2], results["dense"][2])}))


if __name__ == "__main__":
    main()
