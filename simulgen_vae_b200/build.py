"""Builds the engine's shared libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU):
libsimulgen_b200.so (bf16 operands, default) and libsimulgen_b200_fp16.so (the same sources with -DSG_OP16_HALF:
IEEE fp16 operands for the "fp16" precision mode)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsimulgen_b200.so")
OUT_FP16 = os.path.join(HERE, "libsimulgen_b200_fp16.so")
VARIANTS = [(OUT, []), (OUT_FP16, ["-DSG_OP16_HALF"])]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build(out=None):
    outs = [out] if out else [o for o, _ in VARIANTS]
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    for o in outs:
        if not os.path.isfile(o):
            return True
        t = os.path.getmtime(o)
        if any(os.path.getmtime(d) > t for d in deps):
            return True
    return False


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    procs = []
    for out, defs in VARIANTS:
        if not force and not needs_build(out):
            continue
        cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o", out] + defs + sources()
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for out, pr in procs:
        so, se = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(so + se)
            raise RuntimeError("nvcc failed building %s" % out)
        if verbose:
            print(se)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
