#!/bin/bash
# One gpurun call: tests, bench at two batch sizes, per-op step profile, GEMM micro-bench, ncu launch list.
mkdir -p gpurun_out
bash scripts/gpu_check.sh
echo "== bench 32"; timeout 900 python bench.py --batch 32 --steps 6 --warmup 3 > gpurun_out/bench32.json 2> gpurun_out/bench32.err; tail -c 2500 gpurun_out/bench32.json; tail -5 gpurun_out/bench32.err
echo "== bench 64"; timeout 600 python bench.py --batch 64 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench64.json 2> gpurun_out/bench64.err; tail -c 2500 gpurun_out/bench64.json; tail -5 gpurun_out/bench64.err
echo "== profile 32"; timeout 600 python scripts/profile_step.py 32 > gpurun_out/profile_step32.txt 2>&1; head -40 gpurun_out/profile_step32.txt
echo "== gemm bench"; timeout 600 python scripts/gemm_bench.py 32 5 > gpurun_out/gemm_bench32.txt 2>&1; cat gpurun_out/gemm_bench32.txt
