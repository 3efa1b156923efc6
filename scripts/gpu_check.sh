#!/bin/bash
# Runs on the GPU box (via gpurun): kernel unit tests, tcgen05 GEMM tests (separate processes so a
# trap in one cannot poison the others), parity tests, smoke.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels" ; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/kernels.log
for t in test_tc_fprop test_tc_dgrad test_tc_wgrad; do
  echo "== gemm $t"; timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -p no:cacheprovider -k $t 2>&1 | tail -30 | tee gpurun_out/gemm_$t.log
done
echo "== parity" ; timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | tail -60 | tee gpurun_out/parity.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -10 | tee gpurun_out/smoke.log
