#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess.py tests/test_augment.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python scripts/stream_bench.py 32 5 minmax 2>&1 | tee gpurun_out/stream_bench_minmax.txt
