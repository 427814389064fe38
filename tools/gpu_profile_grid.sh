#!/usr/bin/env bash
# ncu --set full of the grid-side kernels of one bench step (fibre passes, generators, theta, band reduce), 1 GPU.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph"
timeout 200 python bench.py $ARGS > "$OUT/plain_grid.log" 2>&1 || { echo "plain run failed"; tail -n 5 "$OUT/plain_grid.log"; exit 1; }
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_fibre_pass|k_b1_gens|k_b1_theta|k_band_reduce' -s 24 -c 8 -f -o "$OUT/prof_grid" \
    python bench.py $ARGS > "$OUT/ncu_full_grid.log" 2>&1
ls -la "$OUT"/prof_grid.ncu-rep
