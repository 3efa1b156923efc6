#!/usr/bin/env bash
# Final single-GPU pass after the static-field work: whole gpu test-suite, smoke(), default bench line, config-4 bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; tail -4 gpurun_out/r2_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -4 gpurun_out/r2_smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_b64_fp16.json 2> gpurun_out/r2_bench_1gpu_b64_fp16.err; echo "bench rc=$?"
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4_compact.json 2> gpurun_out/r2_bench_config4_compact.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu_b64_fp16.json", "gpurun_out/r2_bench_config4_compact.json"):
    d = json.load(open(f))
    print(f, "value %.1f ms %.2f frac %.3f roofline %s %.0f (%.3f) e2e host %.1f resident %s" % (d['value'], d['ms_per_step'], d['step_tensor_frac'],
          d['roofline']['bound'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], (d['e2e']['resident'] or {}).get('value')))
    print("   sweep", json.dumps(d['config']['batch_sweep']))
    print("   cpu", d.get('cpu_baseline'))
PY
