#!/usr/bin/env bash
# B0 (cell-integrated) family iteration on one GPU: its GPU tests, the configs[2] bench line and the warm launch list of a step.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x -k "b0 or B0 or two_live or scan or config" > "$OUT/pytest_b0.log" 2>&1
echo "pytest b0 rc=$? $(tail -n 1 $OUT/pytest_b0.log)"; grep -E "^FAILED|^ERROR|^E  " "$OUT/pytest_b0.log" | head -20
timeout 300 python bench.py --workload b0_cfg3 --no-e2e --no-cpu-baseline > "$OUT/wl_b0_cfg3_plain.json" 2> "$OUT/wl_b0_cfg3_plain.err"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/wl_b0_cfg3_plain.json") if l.startswith("{")][-1])
r = d["roofline"]
print(f"b0_cfg3 ms/step {d['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms  call {r['call_ms']:.4f}  non-K1 {d['ms_per_step'] - r['call_ms']:.4f}  elbo {d['elbo']:.8e} launches/step {d['gpu_launches'] / d['steps']:.0f}")
PY
NLAUNCH=900 bash tools/gpu_launch_list_warm.sh "--workload b0_cfg3" b0
