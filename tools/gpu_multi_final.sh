#!/bin/bash
# The two sharded lines that matter, graph-replayed, with the library's own collective: headline workload and configs[3].
N=${1:-8}
mkdir -p gpurun_out
run() {
    name=$1; shift
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
        bench.py --gpus $N --steps 50 --warmup 5 --no-e2e --no-cpu-baseline "$@" > gpurun_out/final${N}_$name.json 2> gpurun_out/final${N}_$name.err
    rc=$?
    python - "$name" $rc gpurun_out/final${N}_$name.json <<'PY'
import json, sys
name, rc, path = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    r = d["roofline"]
    print(f"{name:20s} rc={rc} ms/step {d['ms_per_step']:.4f}  K1 {r['kernel_ms']:.4f} ms ({r['frac']:.3f} of HBM)  call {r['call_ms']:.4f}  non-K1 {d['ms_per_step'] - r['call_ms']:.4f}  elbo {d['elbo']:.8e}")
except Exception as e:
    print(f"{name:20s} rc={rc} no line ({e})")
PY
    [ $rc -ne 0 ] && tail -5 gpurun_out/final${N}_$name.err
}
run tracks512
run 3d --workload 3d --allreduce peer
