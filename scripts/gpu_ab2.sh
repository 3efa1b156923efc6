#!/bin/bash
# step-level A/B of library builds on one box, alternating order
for lib in "$@"; do
  SIMULGEN_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
done
