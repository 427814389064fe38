#!/usr/bin/env bash
# Whole GPU suite, the default bench line, the graph-replayed bench line and the ncu launch list of one bench step (1 GPU).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > "$OUT/pytest_gpu_all.log" 2>&1
echo "pytest rc=$? $(tail -n 1 $OUT/pytest_gpu_all.log)"
ARGS="--no-e2e --no-cpu-baseline"
for tag in plain graph; do
  extra="--no-cuda-graph"; [ "$tag" = graph ] && extra="--cuda-graph"
  timeout 300 python bench.py $ARGS $extra > "$OUT/bench_$tag.json" 2> "$OUT/bench_$tag.err"
  python - "$OUT/bench_$tag.json" "$tag" <<'PY'
import json, sys
line = None
for l in open(sys.argv[1]):
    if l.startswith("{"):
        line = json.loads(l)
if line is None:
    print(sys.argv[2], "no JSON line")
else:
    r = line["roofline"]
    print(f"{sys.argv[2]:6s} ms/step {line['ms_per_step']:.4f}  K1 kernel {r['kernel_ms']:.4f} ms ({r['frac']:.3f} of HBM peak)  call {r['call_ms']:.4f}  non-K1 {line['ms_per_step'] - r['kernel_ms']:.4f}  elbo {line['elbo']:.6e} launches {line['gpu_launches']}")
PY
done
bash tools/gpu_launch_list_warm.sh
