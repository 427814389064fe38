#!/usr/bin/env bash
# GEMM iteration on one GPU: the dense-path GPU tests, then the configs[2] (B0 family) bench line with its tensor roofline.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x -k "gemm or mode_product or b0 or B0 or dense or chol or two_live or config or svgp or vff" > "$OUT/pytest_gemm.log" 2>&1
echo "pytest gemm rc=$? $(tail -n 1 $OUT/pytest_gemm.log)"; grep -E "^FAILED|^ERROR|^E  " "$OUT/pytest_gemm.log" | head -20
timeout 300 python bench.py --workload b0_cfg3 --no-e2e --no-cpu-baseline > "$OUT/wl_b0_cfg3_gemm.json" 2> "$OUT/wl_b0_cfg3_gemm.err"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/wl_b0_cfg3_gemm.json") if l.startswith("{")][-1])
r = d["roofline"]; t = d.get("tensor_roofline", {})
print(f"b0_cfg3 ms/step {d['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms  elbo {d['elbo']:.8e} tensor {t.get('achieved')} frac {t.get('frac')} per-mode ms {[p['ms'] for p in t.get('per_mode', [])]}")
PY
