#!/bin/bash
# Short-row (Tp == 8, static fields) kernels: parity tests, then the config-4 step profile and (BENCH=1) bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_augment.py -x -q -m gpu -k "pack_unpack or recon or assemble or gn_act" > gpurun_out/short_tests.txt 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/short_tests.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "static or config4 or T1" > gpurun_out/short_parity.txt 2>&1
echo "parity rc=$?"; tail -5 gpurun_out/short_parity.txt
PROFILE_CONFIG=4 timeout 500 python scripts/profile_step.py 512 > gpurun_out/r2_step_profile_config4_b512_short.txt 2>&1
echo "profile rc=$?"; head -24 gpurun_out/r2_step_profile_config4_b512_short.txt
if [ -n "$BENCH" ]; then
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4_short.json 2> gpurun_out/r2_bench_config4_short.err
echo "bench rc=$?"; cat gpurun_out/r2_bench_config4_short.json | head -c 3000
fi
