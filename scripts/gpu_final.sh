#!/bin/bash
# One gpurun call: full GPU suite, smoke, default bench (bf16), fp16 bench, reference arm.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
echo "== bench bf16"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench fp16"; timeout 600 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; tail -c 3000 gpurun_out/bench_fp16.json; tail -3 gpurun_out/bench_fp16.err
