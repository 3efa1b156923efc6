"""Per-op GPU time of one training step (CUDA events around every C-ABI call) + CPU/GPU step time."""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from simulgen_vae_b200 import kernels as K  # noqa: E402
from simulgen_vae_b200.trainer import Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
MODE = sys.argv[2] if len(sys.argv) > 2 else "packed"         # "packed": engine.PackedBatch inputs (bench.py's value loop); "fp32"
cfg = bench.CONFIGS[int(os.environ.get("PROFILE_CONFIG", "2"))]       # BASELINE.json configs[n-1]
dev = torch.device("cuda")
model = bench.build_engine_model(cfg, B, dev)
tr = Trainer(model, lr=1e-3, alpha=1e6)
pool = bench.synthetic_batches(2, B, cfg["num_node"], cfg["num_time"], dev, 1)
import simulgen_vae_b200 as sg  # noqa: E402
from simulgen_vae_b200 import engine  # noqa: E402
if MODE == "packed" and engine.loss_target(cfg["num_time"]) == "operand":
    packed = []
    for x in pool:
        op = torch.empty(1, cfg["num_node"], B, sg.tp_of(cfg["num_time"]), dtype=torch.float16, device=dev)
        K.pack_input(x.contiguous(), op, cfg["num_time"])
        packed.append(engine.PackedBatch(op, cfg["num_time"]))
    pool = packed
print("precision %s, inputs: %s, per-GPU batch %d" % (sg.get_precision(), type(pool[0]).__name__, B))
for i in range(3):
    tr.step(pool[i % 2])
torch.cuda.synchronize()
# CPU-side issue time vs GPU completion time
t0 = time.perf_counter()
tr.step(pool[0])
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print("step: CPU issue %.2f ms, until GPU done %.2f ms" % (t_issue * 1e3, t_total * 1e3))
K.PROFILE_ALL = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.step(pool[1])
e1.record()
torch.cuda.synchronize()
prof = K.PROFILE_ALL
K.PROFILE_ALL = None
agg = collections.defaultdict(lambda: [0, 0.0])
for name, a, b in prof:
    agg[name][0] += 1
    agg[name][1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
print("instrumented step %.2f ms; sum of C-ABI op times %.2f ms over %d calls" % (e0.elapsed_time(e1), tot, len(prof)))
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-24s n=%4d  %8.3f ms  %5.1f%%" % (name, n, ms, 100 * ms / tot))
# biggest individual calls
big = sorted(((a.elapsed_time(b), name, i) for i, (name, a, b) in enumerate(prof)), reverse=True)[:25]
for ms, name, i in big:
    print("  call #%4d %-22s %.3f ms" % (i, name, ms))
# every tensor-core GEMM of the step: time, rate, and the time it loses against 1300 TFLOP/s
K.PROFILE = []
tr.step(pool[0])
torch.cuda.synchronize()
gl = [(a.elapsed_time(b), name, fl) for name, fl, a, b in K.PROFILE]
K.PROFILE = None
print("GEMM calls: %d, %.2f ms, %.1f TFLOP/s average" % (len(gl), sum(g[0] for g in gl), sum(g[2] for g in gl) / sum(g[0] for g in gl) / 1e9))
rows = collections.defaultdict(lambda: [0, 0.0, 0.0])
for ms, name, fl in gl:
    r = rows[(name, round(fl / 1e9, 1))]
    r[0] += 1
    r[1] += ms
    r[2] += fl
for (name, gf), (n, ms, fl) in sorted(rows.items(), key=lambda kv: -(kv[1][1] - kv[1][2] / 1.3e12)):
    print("  %-6s %9.1f GFLOP x%2d  %7.3f ms  %7.1f TFLOP/s  lost %6.3f ms" % (name, gf, n, ms, fl / ms / 1e9, ms - fl / 1.3e12))
