#!/usr/bin/env bash
# One gpurun call (1 GPU): compute-sanitizer memcheck (ONE tool per call, see B200_PROFILING.md) on a small subset of the
# gated tests of the paths that have not run on a B200 yet.  Only worth a call if tools/gpu_check_binned.sh showed a fault.
#   gpurun --timeout 900 -- 'bash tools/gpu_sanitize_new_paths.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VGGP_TEST_UNVERIFIED=1 timeout 800 compute-sanitizer --tool memcheck --error-exitcode 1 \
    python -m pytest tests/test_gpu_new_paths.py -m gpu -x -q \
    -k "(binned_elbo and knots1 and dtype1) or all_outside or (b0_step_scan and knots1 and dtype1) or graphed or fused_metrics" \
    > gpurun_out/sanitizer_new_paths.log 2>&1
echo "rc=$?"; tail -n 25 gpurun_out/sanitizer_new_paths.log
