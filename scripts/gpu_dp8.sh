#!/usr/bin/env bash
# 8-GPU data-parallel check: dp_check in the benched precision, then bench.py with the peer exchange through NVSwitch
# multicast, through plain P2P loads / stores, and (optionally) NCCL.
set -u
N=${1:-8}
TAG=${2:-a}
mkdir -p gpurun_out
OUT=gpurun_out/r2_dp_${N}gpu_$TAG.txt
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "--- dp_check mode=peer (multicast) precision=fp16 world=$N" >> $OUT
SIMULGEN_B200_DP=peer SIMULGEN_B200_DP_MULTICAST=1 timeout 300 $TR scripts/dp_check.py fp16 > gpurun_out/r2_dpcheck_peer_fp16_${N}gpu_$TAG.log 2>&1
echo "rc=$?" >> $OUT
grep -E "dp_check\[|Warning|Error" gpurun_out/r2_dpcheck_peer_fp16_${N}gpu_$TAG.log | tail -4 >> $OUT
run() {  # name extra_args env...
  local name=$1 extra=$2; shift 2
  local line
  line=$(env "$@" timeout 600 $TR bench.py --gpus $N --steps 15 --warmup 3 --no-cpu-baseline --batch-sweep '' $extra 2>gpurun_out/r2_dp_bench_${N}gpu_${name}_$TAG.err | tail -1)
  echo "$line" > gpurun_out/r2_bench_${N}gpu_${name}_$TAG.json
  python - "$name" "$line" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    e = d.get("e2e") or {}
    print("%-16s %2d GPUs  %9.1f samples/s  %7.2f ms/step  gemm %6.1f TFLOP/s share %.3f  sm %s MHz  e2e host %s resident %s  [%s]" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"],
        d["clocks"]["sm_mhz"], e.get("value"), (e.get("resident") or {}).get("value"), d["config"].get("dp_exchange")))
except Exception as ex:
    print("%s FAILED %s" % (sys.argv[1], ex))
PY
}
run peer_multicast --no-e2e SIMULGEN_B200_DP=peer SIMULGEN_B200_DP_MULTICAST=1
run peer_p2p --no-e2e SIMULGEN_B200_DP=peer SIMULGEN_B200_DP_MULTICAST=0
cat $OUT
