#!/bin/bash
# Round-1 evidence pass: streaming micro-bench (all kernels vs the measured HBM peak) and ncu --set full captures of the
# final CTA-pair GEMM (8-warp pipelined epilogue) and of the streaming kernels, on small stand-alone commands.
# Reports are reduced to their raw / details pages on the box (gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_preprocess.py -q -m gpu 2>&1 | tail -2
timeout 600 python scripts/stream_bench.py 32 5 > gpurun_out/stream_bench32.txt 2>&1; cat gpurun_out/stream_bench32.txt
cap() {  # name, kernel regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
}
cap pair_recon_v2 conv_gemm_tc2 5 python scripts/gemm_bench.py 32 1 recon
cap stream_gn "gn_act_fwd|gn_bwd|gn_stats" 9 python scripts/stream_bench.py 16 1 gn_act
cap stream_opt "opt_" 4 python scripts/stream_bench.py 16 1 opt
cap stream_sn "sn_p" 5 python scripts/stream_bench.py 16 1 sn_prepare
cap stream_minmax "minmax" 8 python scripts/stream_bench.py 8 1 minmax
cap stream_assemble "assemble" 2 python scripts/stream_bench.py 16 1 assemble
ls -la gpurun_out | tail -30; du -sh gpurun_out
