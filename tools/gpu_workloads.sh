#!/bin/bash
# One-GPU pass over the bench workloads other than the headline (configs[2] in both families, configs[3] 3-D), device-resident.
mkdir -p gpurun_out
for wl in b1_cfg3 b0_cfg3 3d; do
  for tag in plain graph; do
    extra="--no-cuda-graph"; [ "$tag" = graph ] && extra="--cuda-graph"
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline $extra ${BENCH_EXTRA:-} > gpurun_out/wl_${wl}_$tag.json 2> gpurun_out/wl_${wl}_$tag.err
    rc=$?
    python - $wl $tag $rc <<'PY'
import json, sys
wl, tag, rc = sys.argv[1:4]
try:
    d = json.loads([l for l in open(f"gpurun_out/wl_{wl}_{tag}.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print(f"{wl:8s} {tag:5s} rc={rc} ms/step {d['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms ({r['frac']:.3f} of HBM)  call {r['call_ms']:.4f}  non-K1 {d['ms_per_step'] - r['call_ms']:.4f}  elbo {d['elbo']:.8e} launches/step {d['gpu_launches'] / d['steps']:.0f}")
except Exception as e:
    print(f"{wl:8s} {tag:5s} rc={rc} no line ({e})")
PY
    [ $rc -ne 0 ] && tail -4 gpurun_out/wl_${wl}_$tag.err
  done
done
