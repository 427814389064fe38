"""Turn the raw ncu outputs of a GPU call into the committed summaries under profiles/ (run here, no GPU needed):

    python tools/summarize_profiles.py --round 2 --tag binned_ldg_cap256 \
        --launches gpurun_out/launches_binned_ldg_cap256.csv --report gpurun_out/prof_binned_ldg_cap256.ncu-rep

writes  profiles/r<round>_step_breakdown_<tag>.txt   one bench step, kernel by kernel, with the share of the per-observation kernel
        profiles/r<round>_ncu_summary_<tag>.txt      the metrics that matter per profiled kernel (from --set full)
        profiles/r<round>_k1_traffic_<tag>.json      DRAM bytes per launch of the per-observation kernel
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K1 = re.compile(r"k_obs_b1|k_obs_b0")
PATS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_global_red.sum", "smsp__sass_inst_executed_op_shared"]


def read_launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    out = []
    for r in rows:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit", "ns") in ("us", "usecond"):
                ns *= 1e3
            out.append((r["Kernel Name"], r["Grid Size"], ns))
    return out


def breakdown(launches):
    """The last complete step = the launches between the last two first-kernels of a step (k_b1_gens on the fused B1 grid side,
    k_build_factors on the dense path), K1 included."""
    first = "k_b1_gens" if any("k_b1_gens" in k for k, _, _ in launches) else "k_build_factors"
    starts = [i for i, (k, _, _) in enumerate(launches) if first in k]
    if len(starts) < 2:
        return launches
    steps = [launches[a:b] for a, b in zip(starts[:-1], starts[1:])]
    # bench.py ends with a leg on the round-1 packed kernel (acquisition order): the timed steps are the ones on the default layout
    default = [s for s in steps if any("k_obs_b1_binned" in k or "k_obs_b0s" in k for k, _, _ in s)]
    step = (default or steps)[-1]
    last = [i for i, (k, _, _) in enumerate(step) if "k_b1_theta" in k or "k_bwd_theta" in k]     # a step ends with the theta kernel;
    return step[:last[-1] + 1] if last else step                                               # what follows is setup of the next leg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", type=int, required=True)
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--report")
    args = ap.parse_args()
    prof = os.path.join(ROOT, "profiles")
    if args.launches:
        step = breakdown(read_launches(args.launches))
        vg = [(k, g, ns) for k, g, ns in step if "vggp::" in k]
        tot = sum(ns for _, _, ns in vg)
        k1 = sum(ns for k, _, ns in vg if K1.search(k))
        path = os.path.join(prof, f"r{args.round}_step_breakdown_{args.tag}.txt")
        with open(path, "w") as f:
            f.write("# One bench step, kernel by kernel (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
            f.write(f"# source: {os.path.basename(args.launches)} (last full step of the capture, kernels of libvggp only)\n\n")
            for k, g, ns in vg:
                f.write(f"{ns / 1e3:9.1f} us  {100 * ns / tot:5.1f} %  {g:14s} {k[:110]}\n")
            f.write(f"{tot / 1e3:9.1f} us  total of the step under ncu; share of the per-observation kernel = {100 * k1 / tot:.1f} %\n")
        print("wrote", path)
    if args.report:
        out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        path = os.path.join(prof, f"r{args.round}_ncu_summary_{args.tag}.txt")
        traffic = None
        with open(path, "w") as f:
            for vals in rows[2:]:
                name = vals[ki]
                f.write("---- " + name[:100] + "\n")
                rec = {}
                for h, u, v in zip(hdr, units, vals):
                    if any(p in h for p in PATS) and v not in ("", "no data", "n/a"):
                        f.write(f"  {h} [{u}] = {v}\n")
                    if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                        rec[h] = (v, u)
                if K1.search(name) and "dram__bytes_read.sum" in rec:
                    def to_bytes(vu):
                        v, u = float(vu[0].replace(",", "")), vu[1].lower()
                        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                    traffic = {"kernel": name[:80], "dram_read_bytes": to_bytes(rec["dram__bytes_read.sum"]),
                               "dram_write_bytes": to_bytes(rec["dram__bytes_write.sum"]),
                               "duration": " ".join(rec["gpu__time_duration.sum"]),
                               "source": os.path.basename(args.report) + " (ncu --set full, one launch)"}
        print("wrote", path)
        if traffic:
            tp = os.path.join(prof, f"r{args.round}_k1_traffic_{args.tag}.json")
            json.dump(traffic, open(tp, "w"), indent=1)
            print("wrote", tp)


if __name__ == "__main__":
    main()
