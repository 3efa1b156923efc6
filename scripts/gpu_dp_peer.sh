#!/usr/bin/env bash
# Data parallel on N GPUs of one box: correctness (dp_check: DP == single process on the global batch) and step time of
# the exchange modes - "peer" (sharded optimiser over NVLink peer memory, pipelined against the step; default), the same
# without the pipelining, and "nccl" (bucketed all-reduce overlapped with backward, replicated optimiser).
# Usage: gpurun --gpus N -- bash scripts/gpu_dp_peer.sh N tag [skip_single]
set -u
N=${1:-2}
TAG=${2:-a}
mkdir -p gpurun_out
OUT=gpurun_out/r2_dp_peer_${N}gpu_$TAG.txt
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
check() {  # mode precision [env...]
  local mode=$1 prec=$2; shift 2
  echo "--- dp_check mode=$mode precision=$prec $*" >> $OUT
  env SIMULGEN_B200_DP=$mode "$@" timeout 240 $TR scripts/dp_check.py $prec > gpurun_out/r2_dpcheck_${mode}_${prec}_$TAG.log 2>&1
  echo "rc=$?" >> $OUT
  grep -E "dp_check\[|Warning|Error" gpurun_out/r2_dpcheck_${mode}_${prec}_$TAG.log | tail -4 >> $OUT
}
check peer fp32 A=1
check peer fp16 A=1
check nccl fp16 A=1
fmt() { python - "$1" "$2" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print("%-18s %2d GPUs  %9.1f samples/s  %7.2f ms/step  gemm %6.1f TFLOP/s share %.3f  sm %s MHz  loss %.4g  |grad| %.4g" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"],
        d["clocks"]["sm_mhz"], d["config"]["loss"], d["config"]["grad_norm"]))
except Exception as e:
    print("%-18s FAILED %s" % (sys.argv[1], e))
PY
}
if [ -z "${3:-}" ]; then
  line=$(timeout 400 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep '' 2>/dev/null | tail -1)
  fmt single_gpu "$line"
fi
bench() {  # name [env...]
  local name=$1; shift
  local line
  line=$(env "$@" timeout 500 $TR bench.py --gpus $N --steps 15 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep '' 2>gpurun_out/r2_dp_bench_${name}_$TAG.err | tail -1)
  echo "$line" > gpurun_out/r2_dp_${N}gpu_${name}_$TAG.json
  fmt "$name" "$line"
}
bench dp_peer_multicast SIMULGEN_B200_DP=peer
bench dp_peer_p2p SIMULGEN_B200_DP=peer SIMULGEN_B200_DP_MULTICAST=0
bench dp_nccl SIMULGEN_B200_DP=nccl
cat $OUT
