#!/bin/bash
# A/B of two builds of the library on the same box: GEMM tests, micro-bench, step bench (alternating).
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
for lib in simulgen_vae_b200/_ab/libprev.so simulgen_vae_b200/_ab/lib4.so simulgen_vae_b200/libsimulgen_b200.so; do
  echo "== gemm bench $lib"; SIMULGEN_B200_LIB=$PWD/$lib timeout 600 python scripts/gemm_bench.py 64 5 2>&1 | grep -v "dec.res1\|enc.res0 .*wgrad" | tee gpurun_out/gemm_bench64_$(basename $lib).txt
done
for i in 1 2; do for lib in simulgen_vae_b200/_ab/libprev.so simulgen_vae_b200/_ab/lib4.so simulgen_vae_b200/libsimulgen_b200.so; do
  SIMULGEN_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
done; done
