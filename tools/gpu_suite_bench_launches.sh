#!/usr/bin/env bash
# Whole GPU suite, the default bench line, the graph-replayed bench line and the ncu launch list of one bench step (1 GPU).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > "$OUT/pytest_gpu_all.log" 2>&1
echo "pytest rc=$? $(tail -n 1 $OUT/pytest_gpu_all.log)"
ARGS="--no-e2e --no-cpu-baseline"
for tag in plain graph; do
  extra=""; [ "$tag" = graph ] && extra="--cuda-graph"
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_step.csv" \
    python bench.py --steps 2 --warmup 3 $ARGS > "$OUT/ncu_launches_step.log" 2>&1
python - "$OUT/launches_step.csv" <<'PY'
import csv, io, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
ks = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))) for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
# last occurrence of the binned kernel = last step with the default layout; print the launches from the preceding gens kernel to the theta kernel
idx = [i for i, k in enumerate(ks) if "k_obs_b1_binned" in k[0]]
if idx:
    i = idx[-1]
    a = max(j for j in range(i) if "k_b1_gens" in ks[j][0])
    b = min(j for j in range(i, len(ks)) if "k_b1_theta" in ks[j][0])
    tot = sum(k[2] for k in ks[a:b + 1])
    for k in ks[a:b + 1]:
        print(f"{k[2] / 1e3:9.1f} us  {k[1]:>14s}  {k[0][:100]}")
    print(f"{tot / 1e3:9.1f} us  total")
PY
