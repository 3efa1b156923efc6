#!/bin/bash
# ncu --set full over one step's worth of the streaming (non-GEMM) kernels
mkdir -p gpurun_out
B=${1:-32}
CMD="python bench.py --batch $B --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'recon_|opt_|gn_|pack_input|sn_p' -s 400 -c 140 -f -o gpurun_out/stream_b$B $CMD > gpurun_out/ncu_stream.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_stream.log; ls -la gpurun_out/*.ncu-rep
