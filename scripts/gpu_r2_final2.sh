#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "cuda_graph or trainer or elbo or augment" > gpurun_out/r2_pytest_l.log 2>&1; grep -E "^E  |passed|failed" gpurun_out/r2_pytest_l.log | head -12
timeout 600 python scripts/profile_step.py 64 > gpurun_out/r2_step_profile_l.txt 2>&1; head -5 gpurun_out/r2_step_profile_l.txt
timeout 600 python scripts/profile_step.py 16 > gpurun_out/r2_step_profile16_l.txt 2>&1; head -5 gpurun_out/r2_step_profile16_l.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_full_l.json 2> gpurun_out/r2_bench_full_l.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_full_l.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_full_l.json'))
print("value %.1f ms %.2f frac %.3f" % (d['value'], d['ms_per_step'], d['step_tensor_frac']))
print("e2e host %.1f resident %s" % (d['e2e']['value'], d['e2e']['resident']))
print(json.dumps(d['config']['batch_sweep']))
PY
