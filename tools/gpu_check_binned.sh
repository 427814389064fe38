#!/usr/bin/env bash
# One gpurun call (1 GPU, no ncu) that decides whether the binned K1 path can become the default:
#   gpurun --timeout 1500 -- 'bash tools/gpu_check_binned.sh'
# Every step runs under its own `timeout` (a hung kernel must not eat the call) and writes into gpurun_out/.
# Read gpurun_out/binned_summary.txt afterwards.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
SUM="$OUT/binned_summary.txt"
: > "$SUM"
note() { echo "$*" | tee -a "$SUM"; }

note "== 1. default GPU suite (packed kernel, the verified path)"
timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
note "   rc=$? $(tail -n 1 "$OUT/pytest_default.log")"

note "== 2. binned K1 layout, LDG stream first (VGGP_TEST_STREAMS isolates a hang of the TMA variant)"
VGGP_TEST_UNVERIFIED=1 VGGP_TEST_STREAMS=ldg timeout 600 python -m pytest tests/test_gpu_new_paths.py -m gpu -q -k "binned" > "$OUT/pytest_binned_ldg.log" 2>&1
RC_LDG=$?
note "   ldg rc=$RC_LDG $(tail -n 1 "$OUT/pytest_binned_ldg.log")"
VGGP_TEST_UNVERIFIED=1 VGGP_TEST_STREAMS=tma timeout 600 python -m pytest tests/test_gpu_new_paths.py -m gpu -q -k "binned" > "$OUT/pytest_binned_tma.log" 2>&1
RC_TMA=$?
note "   tma rc=$RC_TMA $(tail -n 1 "$OUT/pytest_binned_tma.log")"
note "== 2b. the other new paths: fused metrics, min-max scaling, B0 scan form (prediction and step)"
VGGP_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_new_paths.py -m gpu -q -k "not binned" > "$OUT/pytest_other_new.log" 2>&1
note "   rc=$? $(tail -n 1 "$OUT/pytest_other_new.log")"

bench() {   # name, extra args...
    local name=$1; shift
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
    local rc=$?
    python - "$OUT/bench_$name.json" "$name" "$rc" <<'PY' | tee -a "$SUM"
import json, sys
path, name, rc = sys.argv[1:4]
line = None
for l in open(path):
    if l.startswith("{"):
        line = json.loads(l)
if line is None:
    print(f"   {name:28s} rc={rc} no JSON line")
else:
    r = line["roofline"]
    print(f"   {name:28s} rc={rc} ms/step {line['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms  "
          f"{r['achieved']:.0f} GB/s  frac {r['frac']:.3f}  elbo {line['elbo']:.6e}")
PY
}

note "== 3. one-process K1 sweep at the headline config (N = 2^26, 512 x 512): K1 ms, step ms, fraction of the HBM peak"
VARIANTS="packed"
[ "$RC_LDG" = 0 ] && VARIANTS="$VARIANTS,binned:ldg:128,binned:ldg:256,binned:ldg:512"
[ "$RC_TMA" = 0 ] && VARIANTS="$VARIANTS,binned:tma:128,binned:tma:256,binned:tma:512"
rm -f "$OUT/sweep_k1.jsonl" "$OUT/sweep_k1_thin.jsonl"
timeout 600 python tools/sweep_k1.py --variants "$VARIANTS" --out "$OUT/sweep_k1.jsonl" 2> "$OUT/sweep_k1.err" | tee -a "$SUM"
note "== 4. thin shard (what one of 8 GPUs sees): N = 2^23"
timeout 300 python tools/sweep_k1.py --n-obs 8388608 --variants "$VARIANTS" --out "$OUT/sweep_k1_thin.jsonl" 2> "$OUT/sweep_k1_thin.err" | tee -a "$SUM"
note "== 5. full bench lines of the default path, plain and graph-replayed"
bench packed
bench packed_graph --cuda-graph
note "done"
