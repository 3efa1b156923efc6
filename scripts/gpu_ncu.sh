#!/bin/bash
# ncu evidence: (1) launch list of a short bench run, (2) full capture of the GEMM kernel on the recon and k5 shapes.
mkdir -p gpurun_out
B=${1:-32}
CMD1="python bench.py --batch $B --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD1 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
CMD2="python scripts/gemm_bench.py $B 1 recon"
$CMD2 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -c 6 -f -o gpurun_out/gemm_recon $CMD2 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm recon capture rc=$?"
CMD3="python scripts/gemm_bench.py $B 1 5120->5120"
$CMD3 > gpurun_out/plain_gemm5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -c 6 -f -o gpurun_out/gemm_k5 $CMD3 > gpurun_out/ncu_gemm5.log 2>&1
echo "gemm k5 capture rc=$?"
ls -la gpurun_out | tail -20
