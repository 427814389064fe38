#!/usr/bin/env bash
# Round 2, call A: ncu on the binned per-observation kernel (LDG stream, run cap 512) at N = 2^26 and at the thin-shard
# size 2^23.  Plain runs first; ncu output under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
for N in 67108864 8388608; do
  ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph --obs-layout binned --binned-stream ldg --run-cap 256 --n-obs $N"
  timeout 200 python bench.py $ARGS > "$OUT/plain_binned_$N.log" 2>&1 || { echo "plain run failed $N"; tail -n 5 "$OUT/plain_binned_$N.log"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_obs_b1_binned -s 3 -c 1 -f -o "$OUT/prof_binned_ldg256_$N" \
      python bench.py $ARGS > "$OUT/ncu_full_binned_$N.log" 2>&1
done
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph --obs-layout binned --binned-stream ldg --run-cap 256"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_binned_ldg256.csv" \
    python bench.py $ARGS > "$OUT/ncu_launches_binned.log" 2>&1
ls -la "$OUT" | tail -n 8
