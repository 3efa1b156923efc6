"""CPU test: libsimulgen_b200.so builds for sm_100a (nvcc cross-compiles without a GPU), loads, and exports
every entry point include/simulgen_b200.h declares; the ctypes table in _lib.py covers the same set.
No compute call is made."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "simulgen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from simulgen_vae_b200 import build, _lib
    build.build()
    names = header_symbols()
    assert len(names) >= 30
    for path in (build.OUT, build.OUT_FP16):          # bf16-operand and fp16-operand builds of the same sources
        lib = ctypes.CDLL(path)
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, (path, missing)
    bound = set(_lib.SIGNATURES) | {"sg_last_error", "sg_version", "sg_device_supported"}
    assert set(names) <= bound, sorted(set(names) - bound)
    assert bound <= set(names), sorted(bound - set(names))
    lib.sg_version.restype = ctypes.c_int
    assert lib.sg_version() >= 1


def test_ctypes_structs_match_header_layout():
    from simulgen_vae_b200 import _lib
    assert ctypes.sizeof(_lib.OptItem) == 8 * 8 + 8 + 6 * 4
    assert ctypes.sizeof(_lib.SnLayer) == 6 * 8 + 2 * 8 + 8 * 4


def test_kernels_refuse_cpu_tensors():
    import pytest
    import torch
    from simulgen_vae_b200 import kernels as K
    with pytest.raises(RuntimeError):
        K.axpy(torch.zeros(4), torch.zeros(4), 1.0, False)
