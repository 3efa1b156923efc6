#!/usr/bin/env bash
# N GPUs: data-parallel Trainer == single process on the full batch for STATIC fields (T = 1, compact path), fp16 and fp32,
# peer and NCCL exchange.  Usage: gpurun --gpus 2 -- bash scripts/gpu_dp_static_check.sh 2
set -u
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/r2_dp_check_static_${N}gpu.txt
for mode in peer nccl; do
  for prec in ${PRECS:-fp16 fp32}; do
    echo "== DP=$mode precision=$prec static (T = 1)" >> gpurun_out/r2_dp_check_static_${N}gpu.txt
    DP_CHECK_STATIC=1 SIMULGEN_B200_DP=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2953$N scripts/dp_check.py $prec 2>&1 | grep -v "OMP_NUM_THREADS\|\*\*\*\*\|Moving model" >> gpurun_out/r2_dp_check_static_${N}gpu.txt
    echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r2_dp_check_static_${N}gpu.txt
  done
done
cat gpurun_out/r2_dp_check_static_${N}gpu.txt
