"""Stand-alone timing of the tcgen05 implicit-GEMM kernel on the headline model's dominant conv shapes
(per-GPU batch B): CUDA-event time, algorithmic TFLOP/s over the valid columns, fraction of the measured
bf16 peak.  Also the command the `ncu --set full` capture of the GEMM kernel is taken on (profiles/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from simulgen_vae_b200 import kernels as K  # noqa: E402
from simulgen_vae_b200.engine import tp_of  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
T = 200
Tp = tp_of(T, "bf16")
dev = torch.device("cuda")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    PEAK = 1400.0

SHAPES = [  # name, Cin, Cout, k
    ("enc.conv0 95008->1024 k1", 95008, 1024, 1),
    ("dec.res2 5120->5120 k5", 5120, 5120, 5),
    ("dec.recon 1024->95008 k1", 1024, 95008, 1),
    ("dec.res2 1024->5120 k1", 1024, 5120, 1),
    ("dec.res1 2560->2560 k5", 2560, 2560, 5),
    ("enc.res0 1024->1024 k3", 1024, 1024, 3),
]


def timed(fn):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(REPS):
        flush.zero_()                       # evict L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name, Cin, Cout, k in SHAPES:
    if ONLY and ONLY not in name:
        continue
    Cin_p = (Cin + 7) // 8 * 8
    wg = (torch.randn(k, Cout, Cin_p, device=dev) * 0.02).to(torch.bfloat16)
    act = torch.randn(k, Cin, B, Tp, device=dev).to(torch.bfloat16)
    act[..., T:] = 0
    dy = torch.randn(k, Cout, B, Tp, device=dev).to(torch.bfloat16)
    dy[..., T:] = 0
    bias = torch.zeros(Cout, device=dev)
    out = torch.empty(Cout, B, Tp, device=dev)
    dx = torch.empty(Cin, B, Tp, device=dev)
    dwg = torch.empty(k, Cout, Cin_p, device=dev)
    flops = 2.0 * Cin * Cout * k * B * T
    G = 8
    stats = torch.empty(B, G, 2, device=dev)
    out16 = torch.empty(Cout, B, Tp, device=dev, dtype=torch.bfloat16) if Cout > 128 else None
    variants = [("fprop", lambda: K.conv_fprop(wg, act, bias, out, Cin)),
                ("fp+gn", lambda: K.conv_fprop_gn(wg, act, bias, out, Cin, stats, T, G))]
    if out16 is not None and "recon" in name:
        variants.append(("gn16", lambda: K.conv_fprop_gn(wg, act, bias, out16, Cin, stats, T, G)))
    for mode, fn in variants + [
                     ("dgrad", lambda: K.conv_dgrad(wg, dy, dx, Cin)),
                     ("wgrad", lambda: K.conv_wgrad(dy, act, dwg, Cin))]:
        ms = timed(fn)
        tf = flops / (ms * 1e-3) / 1e12
        print("%-28s %-5s B=%d  %8.3f ms  %7.1f TFLOP/s  %.3f of measured sustained peak (%.0f)" %
              (name, mode, B, ms, tf, tf / PEAK, PEAK), flush=True)
    del wg, act, dy, out, dx, dwg
    torch.cuda.empty_cache()
