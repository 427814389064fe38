"""One-process A/B sweep of the per-observation kernel variants at the bench configuration (1 GPU): the data set is
generated and the grid forward run once, then every (layout, stream, run_cap) variant is packed, warmed up and timed
with CUDA events -- K1 alone and the whole step.  Much cheaper in GPU minutes than one bench.py process per variant.

    python tools/sweep_k1.py [--n-obs 67108864] [--steps 20] [--out gpurun_out/sweep_k1.jsonl]

Only variants whose gated tests passed should be trusted: run tools/gpu_check_binned.sh first (it calls this)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vggp_b200 as vg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-obs", type=int, default=bench.N_TOTAL)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--variants", default="packed,binned:ldg:128,binned:ldg:256,binned:ldg:512,binned:tma:128,binned:tma:256,binned:tma:512")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_k1.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = vg._lib.load()
    meshes = [torch.linspace(0, 1, k) for k in bench.KNOTS]
    plan = vg.GridPlan(vg.B1_ASVGP, meshes, torch.float32, dev)
    xs, y = bench.make_tracks(0, args.n_obs, args.n_obs, dev, torch.float32)
    theta, m, Ls = bench.make_params(meshes, dev)
    theta_d, m_d = theta.to(dev), m.to(dev)
    L_d = torch.cat([L.reshape(-1) for L in Ls]).to(dev).contiguous()
    peak, peak_src = bench.measured_peaks()
    alg_bytes = args.n_obs * (len(bench.KNOTS) + 1) * 4 + 2 * plan.M * 4
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    ref = None
    with open(args.out, "a") as f:
        for spec in args.variants.split(","):
            parts = spec.split(":")
            try:
                if parts[0] == "packed":
                    obs = plan.pack(xs, y, sort_by_cell=True)
                    extra = {"run_len": obs.run_len}
                else:
                    lib.vggp_set_binned_stream(1 if parts[1] == "tma" else 0)
                    obs = plan.bin(xs, y, run_cap=int(parts[2]))
                    extra = {"n_tasks": obs.n_tasks, "n_runs": obs.n_runs, "streamed_bytes": obs.streamed_bytes}
                for _ in range(args.warmup):
                    out = plan.step(theta_d, m_d, L_d, obs, None)
                torch.cuda.synchronize()
                ea = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
                eb = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for i in range(args.steps):
                    plan.grid_forward(theta_d, m_d, L_d)
                    ea[i].record()
                    plan.obs_fwd_bwd(obs)
                    eb[i].record()
                    out = plan.grid_backward(theta_d, m_d, L_d, 1.0)
                t1.record()
                torch.cuda.synchronize()
                k1 = sum(a.elapsed_time(b) for a, b in zip(ea, eb)) / args.steps
                step = t0.elapsed_time(t1) / args.steps
                elbo = out[0][0].item()
                dm = out[2].clone()
                if ref is None:
                    ref = (elbo, dm)
                rec = {"variant": spec, "k1_ms": k1, "step_ms": step, "k1_gbs": alg_bytes / (k1 * 1e-3) / 1e9,
                       "k1_frac_of_hbm_peak": alg_bytes / (k1 * 1e-3) / 1e9 / peak, "peak": peak, "peak_source": peak_src,
                       "elbo": elbo, "elbo_rel_vs_first": abs(elbo - ref[0]) / abs(ref[0]),
                       "dm_rel_vs_first": ((dm - ref[1]).norm() / ref[1].norm()).item(), "info": plan.read_info(), **extra}
                del obs
            except Exception as e:          # a failing variant must not stop the sweep
                rec = {"variant": spec, "error": repr(e)[:300]}
            finally:
                lib.vggp_set_binned_stream(0)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
