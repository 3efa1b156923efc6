#!/usr/bin/env bash
# Round-2 evidence pass (1 GPU): step profile, streaming micro-benchmark, ncu launch list with DRAM traffic of the very
# bench command, and ncu --set full captures of the dominant GEMM and of the GroupNorm / recon streaming kernels.
# Reports are reduced to raw / details pages on the box (gpurun_out is limited to 64 MiB).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep ''"
timeout 600 python scripts/profile_step.py 64 > gpurun_out/r2_step_profile_b64.txt 2>&1; head -14 gpurun_out/r2_step_profile_b64.txt
timeout 600 python scripts/stream_bench.py 64 5 > gpurun_out/r2_stream_bench_b64.txt 2>&1; cat gpurun_out/r2_stream_bench_b64.txt
eval "$CMD" > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_traffic.csv bash -c "$CMD" > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/r2_launches_traffic.csv gpurun_out/r2_launches_traffic_bench_b64.txt \
    gpurun_out/r2_gemm_traffic_b64.json "$CMD" 64 5 fp16 | head -30
cap() {  # name, kernel regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/r2_ncu_full_${name}_details.txt 2>/dev/null
}
cap pair_k5 conv_gemm_tc2 6 python scripts/gemm_bench.py 64 1 "5120->5120"
cap pair_recon conv_gemm_tc2 6 python scripts/gemm_bench.py 64 1 recon
cap stream_gn "gn_act_fwd|gn_bwd" 12 python scripts/stream_bench.py 64 1 gn_act
cap stream_recon "recon_fwd_fast|recon_bwd_apply_fast" 4 python scripts/stream_bench.py 64 1 recon
rm -f gpurun_out/r2_launches_traffic.csv.tmp
ls -la gpurun_out | grep r2_ | tail -30; du -sh gpurun_out
