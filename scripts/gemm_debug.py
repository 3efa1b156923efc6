"""Debug runner: each (mode, shape) in its own subprocess with CUDA_LAUNCH_BLOCKING=1."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    ("fprop", 24, 72, 1, 3, 21), ("fprop", 200, 72, 1, 3, 21), ("fprop", 24, 300, 1, 3, 21), ("fprop", 24, 72, 1, 30, 21),
    ("fprop", 24, 72, 3, 3, 21), ("fprop", 24, 72, 5, 3, 21),
    ("wgrad", 24, 72, 1, 3, 21), ("wgrad", 24, 300, 1, 3, 21), ("wgrad", 24, 72, 1, 30, 21), ("wgrad", 24, 72, 3, 3, 21),
    ("dgrad", 24, 72, 1, 3, 21), ("dgrad", 24, 200, 1, 3, 21), ("dgrad", 24, 72, 3, 3, 21),
    ("fprop", 256, 384, 3, 4, 200), ("dgrad", 256, 384, 3, 4, 200), ("wgrad", 256, 384, 3, 4, 200),
    ("fprop", 8200, 136, 1, 2, 100), ("fprop", 200, 4104, 1, 2, 100), ("wgrad", 200, 4104, 1, 2, 100),
    ("fprop", 1280, 1280, 5, 2, 200), ("dgrad", 1280, 1280, 5, 2, 200), ("wgrad", 1280, 1280, 5, 2, 200),
]


def child(mode, Cin, Cout, k, B, T):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    import kernel_emulator as emu
    from simulgen_vae_b200 import kernels as K
    from simulgen_vae_b200.engine import tp_of
    from test_gemm_tc_gpu import make
    wg, act, dy, bias, Tp, Cin_p = make(Cin, Cout, k, B, T)
    if mode == "fprop":
        o1 = torch.zeros(Cout, B, Tp, device="cuda"); o2 = torch.empty_like(o1)
        K.conv_fprop(wg, act, bias, o1, Cin); torch.cuda.synchronize()
        emu.conv_fprop(wg, act, bias, o2, Cin)
    elif mode == "dgrad":
        o1 = torch.zeros(Cin, B, Tp, device="cuda"); o2 = torch.empty_like(o1)
        K.conv_dgrad(wg, dy, o1, Cin); torch.cuda.synchronize()
        emu.conv_dgrad(wg, dy, o2, Cin)
    else:
        o1 = torch.zeros(k, Cout, Cin_p, device="cuda"); o2 = torch.empty_like(o1)
        K.conv_wgrad(dy, act, o1, Cin); torch.cuda.synchronize()
        emu.conv_wgrad(dy, act, o2, Cin)
    err = float((o1 - o2).norm() / o2.norm())
    print("RESULT rel_l2=%.3e  max|d|=%.3e" % (err, float((o1 - o2).abs().max())))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1], *[int(a) for a in sys.argv[2:]])
        sys.exit(0)
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    for c in CASES:
        r = subprocess.run([sys.executable, __file__] + [str(a) for a in c], capture_output=True, text=True, env=env,
                           timeout=300)
        out = (r.stdout + r.stderr).strip().splitlines()
        keep = [l for l in out if "RESULT" in l or "rror" in l or "timeout" in l][:4]
        print(c, "rc=%d" % r.returncode, " | ".join(keep) if keep else out[-3:], flush=True)
