#!/usr/bin/env bash
# ncu launch list of one bench step with WARM caches (--cache-control none): per-kernel durations as they are inside a step.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph ${1:-}"
TAG=${2:-warm}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c ${NLAUNCH:-300} --csv --log-file "$OUT/launches_$TAG.csv" \
    python bench.py $ARGS > "$OUT/ncu_launches_$TAG.log" 2>&1
python - "$OUT/launches_$TAG.csv" <<'PY'
import csv, io, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
ks = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))) for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
idx = [i for i, k in enumerate(ks) if "k_obs_b1_binned" in k[0] or "k_obs_b0s" in k[0]]
if idx:
    i = idx[-1]
    first = [j for j in range(i) if "k_b1_gens" in ks[j][0] or "k_build_factors" in ks[j][0]]
    last = [j for j in range(i, len(ks)) if "k_b1_theta" in ks[j][0] or "k_bwd_theta" in ks[j][0]]
    a, b = max(first), min(last)
    tot = sum(k[2] for k in ks[a:b + 1])
    for k in ks[a:b + 1]:
        print(f"{k[2] / 1e3:9.1f} us  {k[1]:>14s}  {k[0][:100]}")
    print(f"{tot / 1e3:9.1f} us  total")
PY
