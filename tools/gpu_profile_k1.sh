#!/usr/bin/env bash
# One gpurun call (1 GPU) with ncu on the per-observation kernel of the chosen layout:
#   gpurun --timeout 1200 -- 'bash tools/gpu_profile_k1.sh binned ldg 256'      (or: packed)
# Plain run first (ncu only after the same command exited 0), then the launch list, then one --set full capture.
set -u
cd "$(dirname "$0")/.."
LAYOUT=${1:-packed}; STREAM=${2:-ldg}; CAP=${3:-256}
OUT=gpurun_out
mkdir -p "$OUT"
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph --obs-layout $LAYOUT --binned-stream $STREAM --run-cap $CAP"
TAG="${LAYOUT}_${STREAM}_cap${CAP}"
python bench.py $ARGS > "$OUT/plain_$TAG.log" 2>&1 || { echo "plain run failed"; tail -n 5 "$OUT/plain_$TAG.log"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_$TAG.csv" \
    python bench.py $ARGS > "$OUT/ncu_launches_$TAG.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_obs_b1 -s 3 -c 2 -o "$OUT/prof_$TAG" \
    python bench.py $ARGS > "$OUT/ncu_full_$TAG.log" 2>&1
ls -la "$OUT" | tail -n 8
