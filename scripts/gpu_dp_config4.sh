#!/usr/bin/env bash
# Pure data parallelism at the static configuration (config 4: 2.29 G parameters) on N GPUs: sharded optimiser over NVLink
# peer memory vs NCCL all-reduce.  Usage: gpurun --gpus N -- 'MODES="peer nccl" bash scripts/gpu_dp_config4.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
for mode in ${MODES:-peer nccl}; do
  SIMULGEN_B200_DP=$mode timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --config 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/r2_bench_config4_${N}gpu_$mode.err | grep "^{" > gpurun_out/r2_bench_config4_${N}gpu_$mode.json
  echo "$mode rc=${PIPESTATUS[0]}"; tail -2 gpurun_out/r2_bench_config4_${N}gpu_$mode.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_config4_${N}gpu_$mode.json"))
    print("$mode N=$N value %.1f ms %.2f dp %s" % (d["value"], d["ms_per_step"], json.dumps(d["config"].get("dp_exchange"))[:400]))
except Exception as e:
    print("no line:", e)
PY
done
