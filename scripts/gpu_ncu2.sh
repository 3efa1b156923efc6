#!/bin/bash
# launch list with DRAM traffic per launch (3 metrics, single pass) of a short bench run
mkdir -p gpurun_out
B=${1:-64}
CMD="python bench.py --batch $B --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_traffic.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/launches_traffic.csv)"; tail -2 gpurun_out/ncu_bench.log
