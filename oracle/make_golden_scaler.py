"""TEST INFRASTRUCTURE ONLY - golden output of the reference's preprocessing scan.

Runs the UNMODIFIED /root/reference/modules/data_preprocess.py::data_scaler (sklearn MinMaxScaler fitted on the seeded
row sample, chunked in-place transform, scaler.pkl) on small synthetic [P, T, N] datasets - float64 like the reference's
np.zeros-built arrays and float32 - including a constant node and a node with a NaN, and records the inputs' generator
seed, the fitted scaler attributes, the sampled row indices and the scaled field.
Output: tests/golden/scaler_toy.npz.  Usage: python oracle/make_golden_scaler.py"""
import contextlib
import importlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

CASES = dict(f64=(np.float64, 30, 40, 23), f32=(np.float32, 5, 24, 50))     # dtype, P, T, N


def make_data(dtype, P, T, N, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(P, T, N)) * rng.uniform(0.1, 30.0, size=N) + rng.uniform(-5, 5, size=N)
    x[:, :, 3] = 1.25                       # constant node: data_range 0 -> scale handled by _handle_zeros_in_scale
    x = x.astype(dtype)
    x[1, 2, 5] = np.nan                     # MinMaxScaler ignores NaNs in fit and keeps them in transform
    return x


def load_reference_module():
    ref_import._install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    try:
        return importlib.import_module("modules.data_preprocess")
    finally:
        sys.path.remove(ref_import.REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def main():
    ref = load_reference_module()
    out = {}
    cwd = os.getcwd()
    for i, (name, (dtype, P, T, N)) in enumerate(CASES.items()):
        x = make_data(dtype, P, T, N, 100 + i)
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "model_save"))
            os.chdir(tmp)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    y, shape, scaler = ref.data_scaler(x.copy(), x, T, N, 1, chunk_size=64)
            finally:
                os.chdir(cwd)
        assert y.dtype == dtype and tuple(shape) == (T, N)
        out[name + "_seed"] = np.int64(100 + i)
        out[name + "_shape"] = np.array([P, T, N])
        out[name + "_scaled"] = y
        for attr in ("data_min_", "data_max_", "data_range_", "scale_", "min_"):
            out[name + "_" + attr] = getattr(scaler, attr)
        out[name + "_n_samples_seen"] = np.int64(scaler.n_samples_seen_)
        print(name, "scaled range", np.nanmin(y), np.nanmax(y), "samples seen", scaler.n_samples_seen_, y.dtype, scaler.scale_.dtype)
    path = os.path.join(ROOT, "tests", "golden", "scaler_toy.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
