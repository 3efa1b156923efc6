#!/usr/bin/env bash
# Data-parallel experiments on N GPUs of one box (N = $1, default 2): correctness (dp_check) and the step-time cost of the
# overlapped gradient all-reduce under different SM reservations / NCCL CTA limits.  Usage: gpurun --gpus N -- bash scripts/gpu_dp_sweep.sh N tag
set -u
N=${1:-2}
TAG=${2:-a}
mkdir -p gpurun_out
OUT=gpurun_out/r2_dp_sweep_${N}gpu_$TAG.txt
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for prec in fp32 fp16; do
  SIMULGEN_B200_CHUNK_WGRAD_MELEMS=0.1 SIMULGEN_B200_CHUNK_WGRAD_ALIGN=64 timeout 300 $TR scripts/dp_check.py $prec 2>&1 | grep -E "dp_check|Error|error" >> $OUT
done
run() {   # name, env...
  local name=$1; shift
  local line
  line=$(env "$@" timeout 400 $TR bench.py --gpus $N --steps 15 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep '' 2>gpurun_out/r2_dp_${name}.err | tail -1)
  echo "$line" > gpurun_out/r2_dp_${N}gpu_${name}_$TAG.json
  python - "$name" "$line" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print("%-28s %2d GPUs  %8.1f samples/s  %7.2f ms/step  gemm %6.1f TFLOP/s share %.3f  sm %s MHz" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("%-28s FAILED %s" % (sys.argv[1], e))
PY
}
# single GPU on this box first (same binary, same clocks regime)
line=$(timeout 400 python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep '' 2>/dev/null | tail -1)
python - "single_gpu" "$line" >> $OUT <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print("%-28s %2d GPUs  %8.1f samples/s  %7.2f ms/step  gemm %6.1f TFLOP/s share %.3f  sm %s MHz" % (
    sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["clocks"]["sm_mhz"]))
PY
run default_reserve8_chunk4 A=1
run r1_behaviour_reserve0_chunk1 SIMULGEN_B200_DP_RESERVE_SMS=0 SIMULGEN_B200_CHUNK_WGRAD_PARTS=1
run reserve0_chunk4 SIMULGEN_B200_DP_RESERVE_SMS=0
run reserve16_chunk4 SIMULGEN_B200_DP_RESERVE_SMS=16
run reserve8_maxctas8 NCCL_MAX_CTAS=8
run reserve16_maxctas16 SIMULGEN_B200_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16
run no_overlap SIMULGEN_B200_DP_OVERLAP=0
cat $OUT
