#!/usr/bin/env bash
# BASELINE configs 3, 4, 5 at real size: full-size parity cases + a short bench line each.
set -u
TAG=${1:-a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "full_size" > gpurun_out/r2_pytest_fullsize_$TAG.log 2>&1
tail -15 gpurun_out/r2_pytest_fullsize_$TAG.log
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config${c}_$TAG.json 2> gpurun_out/r2_bench_config${c}_$TAG.err
  echo "config $c rc=$?"; tail -3 gpurun_out/r2_bench_config${c}_$TAG.err; cut -c1-600 gpurun_out/r2_bench_config${c}_$TAG.json
done
