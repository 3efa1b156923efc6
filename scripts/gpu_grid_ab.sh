#!/bin/bash
# Grid sizing of the persistent streaming kernels: short-row kernels sized by occupancy (config 4) and the
# GroupNorm kernels' task grid (SIMULGEN_B200_GN_GRID=0: fixed 148 x 8 blocks, 1: resident grid for many-task layers).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_augment.py -x -q -m gpu -k "pack_unpack or recon or assemble or gn_act" > gpurun_out/short_tests.txt 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/short_tests.txt
PROFILE_CONFIG=4 timeout 500 python scripts/profile_step.py 512 > gpurun_out/r2_step_profile_config4_b512_short.txt 2>&1
echo "profile rc=$?"; head -16 gpurun_out/r2_step_profile_config4_b512_short.txt
for mode in 0 1 0 1; do
  SIMULGEN_B200_GN_GRID=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --batch-sweep '' > gpurun_out/grid_ab_bench_$mode.json 2> gpurun_out/grid_ab_bench_$mode.err
  python - <<PY
import json
d = json.load(open("gpurun_out/grid_ab_bench_$mode.json"))
print("GN_GRID=$mode value %.1f ms %.3f frac %.4f clk %s" % (d["value"], d["ms_per_step"], d["step_tensor_frac"], d["clocks"]["sm_mhz"]))
PY
done
for mode in 0 1; do
  SIMULGEN_B200_GN_GRID=$mode timeout 600 python scripts/stream_bench.py 64 5 gn_act > gpurun_out/grid_ab_stream_$mode.txt 2>&1
  echo "== stream GN_GRID=$mode"; tail -25 gpurun_out/grid_ab_stream_$mode.txt
done
