#!/bin/bash
# ncu source-level (SASS + stall sampling) export of the GroupNorm backward kernels on a stand-alone command
mkdir -p gpurun_out
CMD="python scripts/stream_bench.py 32 1 gn_act"
$CMD > gpurun_out/plain_gn.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_bwd_reduce|gn_bwd_apply" -s 2 -c 2 -f -o /tmp/gnsrc $CMD > gpurun_out/ncu_gnsrc.log 2>&1
echo "rc=$?"
ncu -i /tmp/gnsrc.ncu-rep --page source --csv > gpurun_out/gn_bwd_source.csv 2>/dev/null
ncu -i /tmp/gnsrc.ncu-rep --page details > gpurun_out/gn_bwd_details.txt 2>/dev/null
ls -la gpurun_out/gn_bwd_source.csv gpurun_out/gn_bwd_details.txt
