#!/usr/bin/env bash
# Round-2 GPU check: the gpu test-suite, the streaming-kernel micro-benchmark of the GroupNorm kernels and a short bench.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_r2_check.sh <tag> [pytest -k expression]
set -u
TAG=${1:-a}
KEXPR=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu_$TAG.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -k "$KEXPR" > gpurun_out/r2_pytest_$TAG.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/r2_pytest_$TAG.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/r2_pytest_$TAG.log
tail -5 gpurun_out/r2_pytest_$TAG.log
timeout 300 python scripts/stream_bench.py 64 5 gn_act > gpurun_out/r2_stream_gn_$TAG.txt 2>&1
cat gpurun_out/r2_stream_gn_$TAG.txt
timeout 600 python bench.py --precision fp16 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_fp16_$TAG.json 2> gpurun_out/r2_bench_fp16_$TAG.err
cat gpurun_out/r2_bench_fp16_$TAG.json
timeout 600 python scripts/profile_step.py 64 > gpurun_out/r2_step_profile_$TAG.txt 2>&1
head -30 gpurun_out/r2_step_profile_$TAG.txt
