#!/bin/bash
# Multi-GPU check (run with `gpurun --gpus N -- bash tools/gpu_multi.sh N`): the two-rank tests of the peer-memory all-reduce,
# then the sharded bench step with NCCL / the library's own collective, eager / graph-replayed, time- / cell-range shards.
# Every command has its own timeout (a missing rank shows up as a bounded barrier timeout, not a hang).
N=${1:-2}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_multi.log)"
grep -h "multicast path" gpurun_out/pytest_multi.log | head -2
fi
run() {   # name, extra args
    name=$1; shift
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
        bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --no-e2e --no-cpu-baseline "$@" > gpurun_out/multi${N}_$name.json 2> gpurun_out/multi${N}_$name.err
    rc=$?
    python - "$name" $rc gpurun_out/multi${N}_$name.json <<'PY'
import json, sys
name, rc, path = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    r = d["roofline"]
    print(f"{name:28s} rc={rc} ms/step {d['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms ({r['frac']:.3f} of HBM)  call {r['call_ms']:.4f}  non-K1 {d['ms_per_step'] - r['call_ms']:.4f}  elbo {d['elbo']:.8e}  n/gpu {d['config']['n_obs_per_gpu']}")
except Exception as e:
    print(f"{name:28s} rc={rc} no line ({e})")
PY
    [ $rc -ne 0 ] && tail -5 gpurun_out/multi${N}_$name.err
}
if [ "${RUNSET:-full}" = "full" ]; then
run nccl --allreduce nccl --no-cuda-graph
run nccl_graph --allreduce nccl --cuda-graph
run peer --allreduce peer --no-cuda-graph
run peer_graph --allreduce peer --cuda-graph
run peer_graph_reshard --allreduce peer --cuda-graph --spatial-reshard
run nccl_reshard --allreduce nccl --spatial-reshard --no-cuda-graph
elif [ "${RUNSET}" = "cap" ]; then      # run-cap check: automatic against 256, time- and cell-range shards
run auto --allreduce ${AR:-nccl}
run cap256 --allreduce ${AR:-nccl} --run-cap 256
run auto_reshard --allreduce ${AR:-nccl} --spatial-reshard
run cap256_reshard --allreduce ${AR:-nccl} --spatial-reshard --run-cap 256
else      # the short set for a large-N call (every line replays CUDA graphs, the bench default)
run nccl_graph --allreduce nccl
run nccl_graph_reshard --allreduce nccl --spatial-reshard
run peer_graph --allreduce peer
run peer_graph_reshard --allreduce peer --spatial-reshard
run 3d_nccl_graph --workload 3d --allreduce nccl
run 3d_peer_graph --workload 3d --allreduce peer
fi
