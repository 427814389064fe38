#!/usr/bin/env bash
# BASELINE-config parity tests + the default bench line (1 GPU).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q --durations=12 > "$OUT/pytest_configs.log" 2>&1
echo "configs rc=$? $(tail -n 1 $OUT/pytest_configs.log)"
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_configs.py > "$OUT/pytest_gpu_rest.log" 2>&1
echo "rest rc=$? $(tail -n 1 $OUT/pytest_gpu_rest.log)"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
echo "bench rc=$?"; tail -c 3000 "$OUT/bench_default.json"
