#!/bin/bash
# Round-end run on one B200: full GPU suite, smoke, default bench (bf16) and the reference arm, fp16 bench, ncu launch list.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep smoke
echo "== bench bf16"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2600 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
echo "== bench fp16"; timeout 600 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; tail -c 700 gpurun_out/bench_fp16.json
echo "== step profile"; timeout 300 python scripts/profile_step.py 64 > gpurun_out/profile_step64.txt 2>&1; head -14 gpurun_out/profile_step64.txt
echo "== launch list"
CMD="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_traffic.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/launches_traffic.csv)"
python scripts/summarize_launches.py gpurun_out/launches_traffic.csv gpurun_out/launches_traffic.txt gpurun_out/gemm_traffic.json "$CMD" 64 5 | head -20
