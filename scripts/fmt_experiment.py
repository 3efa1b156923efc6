"""Experiment: does tcgen05.mma kind::f16 accept DIFFERENT formats for A and B (fp16 x bf16)?
Operands are created as fp16 bit patterns viewed as bf16 tensors; SG_TC_FMT selects the formats in the descriptor."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from simulgen_vae_b200 import kernels as K
import kernel_emulator as emu
fmt = os.environ.get("SG_TC_FMT", "11")
Cin, Cout, k, B, T = 256, 384, 3, 4, 200
Tp = 200
dev = "cuda"
g = torch.Generator().manual_seed(0)
w = (torch.randn(k, Cout, Cin, generator=g) * 0.05).to(dev)
a = torch.randn(Cin, B, T, generator=g).to(dev)
dy = torch.randn(Cout, B, T, generator=g).to(dev)
def as_fmt(t, f):   # f: '0' fp16, '1' bf16 -> tensor typed bf16 holding the chosen bit pattern, plus its fp32 value
    if f == "0":
        h = t.to(torch.float16)
        return h.view(torch.bfloat16), h.float()
    h = t.to(torch.bfloat16)
    return h, h.float()
def planes(val, P, f):
    out = torch.empty(P, val.shape[0], B, Tp, device=dev, dtype=torch.float32)
    emu.write_planes(out, val, T)
    bits, vals = as_fmt(out, f)
    return bits.contiguous(), vals
# fprop: A = W, B = act
wb, wv = as_fmt(w, fmt[0]); ab, av = planes(a, k, fmt[1])
o1 = torch.empty(Cout, B, Tp, device=dev); o2 = torch.empty_like(o1)
K.conv_fprop(wb.contiguous(), ab, None, o1, Cin)
emu.conv_fprop(wv, av, None, o2, Cin)
torch.cuda.synchronize()
print("fmt", fmt, "fprop rel err", float((o1 - o2).norm() / o2.norm()))
# dgrad: A = W, B = dy
db, dv = planes(dy, k, fmt[1])
d1 = torch.empty(Cin, B, Tp, device=dev); d2 = torch.empty_like(d1)
K.conv_dgrad(wb.contiguous(), db, d1, Cin)
emu.conv_dgrad(wv, dv, d2, Cin)
print("fmt", fmt, "dgrad rel err", float((d1 - d2).norm() / d2.norm()))
# wgrad: A = dy, B = act
db2, dv2 = planes(dy, k, fmt[0]); ab2, av2 = planes(a, k, fmt[1])
g1 = torch.empty(k, Cout, Cin, device=dev); g2 = torch.empty_like(g1)
K.conv_wgrad(db2, ab2, g1, Cin)
emu.conv_wgrad(dv2, av2, g2, Cin)
print("fmt", fmt, "wgrad rel err", float((g1 - g2).norm() / g2.norm()))
