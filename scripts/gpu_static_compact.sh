#!/bin/bash
# Static fields (T = 1): compact [C][B] path of the two N-channel layers - kernel tests, engine parity, config-4 profile / bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "static or rows_compact" > gpurun_out/static_tests.txt 2>&1
echo "kernel tests rc=$?"; tail -4 gpurun_out/static_tests.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -rs -m gpu -k "static or config4" > gpurun_out/static_parity.txt 2>&1
echo "parity rc=$?"; tail -9 gpurun_out/static_parity.txt
PROFILE_CONFIG=4 timeout 500 python scripts/profile_step.py 512 > gpurun_out/r2_step_profile_config4_b512_compact.txt 2>&1
echo "profile rc=$?"; head -26 gpurun_out/r2_step_profile_config4_b512_compact.txt
if [ -n "$BENCH" ]; then
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4_compact.json 2> gpurun_out/r2_bench_config4_compact.err
echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench_config4_compact.json; tail -3 gpurun_out/r2_bench_config4_compact.err
fi
