#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "cuda_graph or head or philox or recon or peer" > gpurun_out/r2_pytest_j.log 2>&1; tail -12 gpurun_out/r2_pytest_j.log
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4_b.json 2> gpurun_out/r2_bench_config4_b.err; echo "config4 rc=$?"; tail -3 gpurun_out/r2_bench_config4_b.err; cut -c1-400 gpurun_out/r2_bench_config4_b.json
timeout 900 python bench.py > gpurun_out/r2_bench_full_j.json 2> gpurun_out/r2_bench_full_j.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_full_j.err
