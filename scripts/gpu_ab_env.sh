#!/bin/bash
# step-level A/B of an environment switch on one box, alternating order: gpu_ab_env.sh VAR v1 v2 v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$var=$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
done
