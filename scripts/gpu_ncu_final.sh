#!/bin/bash
# ncu --set full captures of the final kernels on small stand-alone commands (fast replays).  The reports are reduced to
# their raw / details pages on the box (gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out
cap() {  # name, kernel regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
}
cap pair_recon conv_gemm_tc2 6 python scripts/gemm_bench.py 32 1 recon
cap pair_k5 conv_gemm_tc2 6 python scripts/gemm_bench.py 32 1 "5120->5120"
cap stream_recon "recon_fwd_fast|recon_bwd_apply_fast" 6 python scripts/stream_bench.py 16 1 recon
ls -la gpurun_out | tail -12
