#!/usr/bin/env bash
# Final single-GPU pass of round 2: the whole gpu test-suite, smoke(), the default bench line and the reference arm.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; tail -4 gpurun_out/r2_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -4 gpurun_out/r2_smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_b64_fp16.json 2> gpurun_out/r2_bench_1gpu_b64_fp16.err; echo "bench rc=$?"
timeout 900 python bench.py --precision bf16 --no-cpu-baseline --batch-sweep '' > gpurun_out/r2_bench_1gpu_b64_bf16.json 2> gpurun_out/r2_bench_1gpu_b64_bf16.err; echo "bench bf16 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_cpu.json 2> gpurun_out/r2_bench_reference_cpu.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_reference_cpu.json | cut -c1-400
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu_b64_fp16.json", "gpurun_out/r2_bench_1gpu_b64_bf16.json"):
    d = json.load(open(f))
    print(f, "value %.1f ms %.2f frac %.3f gemm %.0f (%.3f) e2e host %.1f resident %s" % (d['value'], d['ms_per_step'], d['step_tensor_frac'],
          d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], (d['e2e']['resident'] or {}).get('value')))
    print("   sweep", json.dumps(d['config']['batch_sweep']))
    print("   cpu", d.get('cpu_baseline'))
PY
