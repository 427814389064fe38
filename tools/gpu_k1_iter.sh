#!/usr/bin/env bash
# One short gpurun call of the K1 tuning loop: binned GPU tests, then the one-process sweep at N = 2^26 and 2^23.
#   gpurun --timeout 600 -- 'bash tools/gpu_k1_iter.sh [variants]'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
VARIANTS=${1:-packed,binned:ldg:128,binned:ldg:256,binned:ldg:512,binned:tma:256,binned:tma:512}
timeout 300 python -m pytest tests/test_gpu_new_paths.py -m gpu -x -q -k "binned" > "$OUT/iter_pytest.log" 2>&1
echo "pytest rc=$? $(tail -n 1 $OUT/iter_pytest.log)"
rm -f "$OUT/iter_sweep.jsonl" "$OUT/iter_sweep_thin.jsonl"
timeout 300 python tools/sweep_k1.py --variants "$VARIANTS" --out "$OUT/iter_sweep.jsonl" 2> "$OUT/iter_sweep.err" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('2^26', r['variant'], r.get('error') or 'k1 %.4f ms step %.4f frac %.3f elbo_rel %.1e dm_rel %.1e' % (r['k1_ms'], r['step_ms'], r['k1_frac_of_hbm_peak'], r['elbo_rel_vs_first'], r['dm_rel_vs_first']))
"
timeout 300 python tools/sweep_k1.py --n-obs 8388608 --variants "$VARIANTS" --out "$OUT/iter_sweep_thin.jsonl" 2> "$OUT/iter_sweep_thin.err" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('2^23', r['variant'], r.get('error') or 'k1 %.4f ms step %.4f frac %.3f elbo_rel %.1e dm_rel %.1e' % (r['k1_ms'], r['step_ms'], r['k1_frac_of_hbm_peak'], r['elbo_rel_vs_first'], r['dm_rel_vs_first']))
"
