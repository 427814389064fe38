#!/usr/bin/env bash
# One 1-GPU call that re-establishes the measured state of the tree: whole GPU suite, the default bench line (with e2e and
# cpu_baseline), the graph-replayed line, the other workloads, the warm and cold ncu launch lists of a bench step and
# one --set full capture of the per-observation kernel and of the grid-side kernels.  Everything lands in gpurun_out/;
# tools/summarize_profiles.py turns it into profiles/ afterwards.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > "$OUT/state_gpu.txt" 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > "$OUT/pytest_gpu_all.log" 2>&1
echo "pytest rc=$? $(tail -n 1 $OUT/pytest_gpu_all.log)"
grep -E "FAILED|ERROR" "$OUT/pytest_gpu_all.log" | head -20
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$? $(tail -n 1 $OUT/smoke.log)"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench default rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench reference rc=$?"
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-cuda-graph > "$OUT/bench_eager.json" 2> "$OUT/bench_eager.err"; echo "bench eager rc=$?"
python - <<'PY'
import json
for tag in ("default", "eager", "reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{tag}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f"{tag:9s} value {d.get('value'):.4e} ms/step {d.get('ms_per_step'):.4f} K1 {r.get('kernel_ms')} frac {r.get('frac')} e2e {d.get('e2e', {}).get('value')} launches {d.get('gpu_launches')} elbo {d.get('elbo')}")
    except Exception as e:
        print(tag, "no line", e)
PY
bash tools/gpu_workloads.sh
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cuda-graph"
# the launch list of the bench command itself (graph replays: ncu lists the kernel nodes one by one), then the eager variants
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_cold.csv" \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launches_cold.log" 2>&1; echo "cold launch list rc=$?"
bash tools/gpu_launch_list_warm.sh
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_obs_b1 -s 3 -c 1 -f -o "$OUT/prof_k1" \
    python bench.py $ARGS > "$OUT/ncu_full_k1.log" 2>&1; echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on \
    -k regex:'k_fibre_pass|k_b1_gens|k_b1_theta|k_band_reduce' -s 24 -c 8 -f -o "$OUT/prof_grid" \
    python bench.py $ARGS > "$OUT/ncu_full_grid.log" 2>&1; echo "ncu grid rc=$?"
ls -la "$OUT" | tail -n 30
