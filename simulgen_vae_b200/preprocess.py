"""SURVEY §8f N4 - the preprocessing scan of the reference on the GPU.

Reference: `data_scaler()` (modules/data_preprocess.py:65-165) fits `MinMaxScaler(feature_range=(-0.7, 0.7))` on a
seeded random sample of the rows of the `[P*T, N]` field matrix, transforms the whole matrix chunk by chunk in place
(`X *= scale_; X += min_`), pickles the scaler to ./model_save/scaler.pkl, and SimulGen-VAE.py:281-283 then transposes
to `[P, N, T]` and casts to float32.

Here the per-node min/max and the affine transform are two HBM-streaming CUDA kernels (csrc/minmax.cu) working in
the dtype of the data like sklearn does, so the scaler attributes and the scaled field are bit-identical; the scaler
object handed back (and pickled) is a real `sklearn.preprocessing.MinMaxScaler` whose `scale_` / `min_` are computed
by sklearn itself from the two rows (data_min, data_max).

* `data_scaler(...)`: the reference's signature and return values (host numpy in, the same array scaled in place out).
* `data_scaler_to_device(...)`: the B200-first entry - returns the float32 `[P, N, T]` training tensor resident in HBM
  (what `create_augmented_dataloaders(..., load_all=True)` moves to the GPU anyway), written by the same sweep that
  applies the transform, without the host round trip.
There is no CPU fallback: both need the CUDA extension."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import kernels as K

FEATURE_RANGE = (-0.7, 0.7)              # data_preprocess.py:90


def reference_sample_rows(total_samples: int) -> np.ndarray:
    """Row indices the reference fits its scaler on (data_preprocess.py:93-108): 10 % of the rows, at most 50 000, at
    least 1000 (or all of them), drawn without replacement after `np.random.seed(42)`.  Like the reference this
    reseeds NumPy's global generator."""
    max_samples = min(50000, total_samples // 10)
    if max_samples < 1000:
        max_samples = min(1000, total_samples)
    np.random.seed(42)
    if total_samples > max_samples:
        return np.random.choice(total_samples, max_samples, replace=False)
    return np.arange(total_samples)


def make_scaler(data_min: np.ndarray, data_max: np.ndarray, n_samples: int, feature_range=FEATURE_RANGE):
    """A fitted sklearn MinMaxScaler with the given per-feature extrema: sklearn derives data_range_, scale_ and min_
    itself from the two-row matrix (data_min, data_max), in the dtype of the data, exactly as MinMaxScaler.fit does on
    the full sample (min and max of a set do not depend on how many other rows there are)."""
    try:
        from sklearn.preprocessing import MinMaxScaler
    except ImportError as e:  # pragma: no cover
        raise ImportError("simulgen_b200.preprocess needs scikit-learn for scaler.pkl compatibility "
                          "(the reference requires it as well)") from e
    sc = MinMaxScaler(feature_range=feature_range).fit(np.stack([np.asarray(data_min), np.asarray(data_max)]))
    sc.n_samples_seen_ = int(n_samples)
    return sc


def _device(device):
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError("simulgen_b200: the preprocessing scan runs on a CUDA device (there is no CPU fallback)")
    return dev


def fit_minmax(x: torch.Tensor, rows: torch.Tensor = None, state=None):
    """x [R, N] on the device; rows int64 indices or None.  Returns (data_min, data_max) device tensors; pass the
    previous result as `state` to extend it with another chunk (partial_fit semantics)."""
    N = x.shape[1]
    if state is None:
        mn = torch.empty(N, dtype=x.dtype, device=x.device)
        mx = torch.empty(N, dtype=x.dtype, device=x.device)
        K.minmax_fit(x, rows, mn, mx, merge=False)
        return mn, mx
    K.minmax_fit(x, rows, state[0], state[1], merge=True)
    return state


def _scaler_vectors(scaler, dtype, dev):
    return (torch.from_numpy(np.ascontiguousarray(scaler.scale_)).to(dev, dtype),
            torch.from_numpy(np.ascontiguousarray(scaler.min_)).to(dev, dtype))


def _rows_per_chunk(total_rows, row_bytes, num_time, chunk_bytes):
    r = max(1, int(chunk_bytes // row_bytes))
    if r >= num_time:
        r = r // num_time * num_time            # whole parameter sets, so that a chunk can emit [P, N, T]
    return min(total_rows, r)


def _scan(flat: np.ndarray, num_time: int, dev, want_host: bool, want_device_t: bool, chunk_bytes: int,
          resident_bytes: int):
    """Shared two-pass driver: fit on the reference's sampled rows, then transform.  Returns (scaler, out_t)."""
    total, N = flat.shape
    if flat.dtype not in (np.float64, np.float32):
        raise TypeError("simulgen_b200.preprocess: float64 or float32 data expected, got %s" % flat.dtype)
    tdtype = torch.float64 if flat.dtype == np.float64 else torch.float32
    sample = np.sort(reference_sample_rows(total))
    row_bytes = N * flat.dtype.itemsize
    rpc = _rows_per_chunk(total, row_bytes, num_time, chunk_bytes)
    bounds = [(a, min(total, a + rpc)) for a in range(0, total, rpc)]
    keep = total * row_bytes <= resident_bytes
    resident = {}
    state = None
    # pass 1: per-node extrema over the sampled rows
    for a, b in bounds:
        lo, hi = np.searchsorted(sample, a), np.searchsorted(sample, b)
        if keep:
            chunk = torch.from_numpy(flat[a:b]).to(dev)
            resident[a] = chunk
            if hi > lo:
                rows = torch.from_numpy(sample[lo:hi] - a).to(dev)
                state = fit_minmax(chunk, rows, state)
        elif hi > lo:
            picked = torch.from_numpy(flat[sample[lo:hi]]).to(dev)      # only the sampled rows travel in this pass
            state = fit_minmax(picked, None, state)
    data_min, data_max = state[0].cpu().numpy(), state[1].cpu().numpy()
    scaler = make_scaler(data_min, data_max, len(sample))
    scale, minv = _scaler_vectors(scaler, tdtype, dev)
    # pass 2: transform (in place on the host array and / or into the float32 [P, N, T] device tensor)
    out_t = None
    if want_device_t:
        if total % num_time != 0 or (rpc % num_time != 0 and rpc != total):
            raise ValueError("simulgen_b200.preprocess: the [P, N, T] output needs chunks of whole parameter sets "
                             "(raise chunk_bytes above one parameter set: %d bytes)" % (row_bytes * num_time))
        out_t = torch.empty(total // num_time, N, num_time, dtype=torch.float32, device=dev)
    for a, b in bounds:
        chunk = resident.pop(a) if keep else torch.from_numpy(flat[a:b]).to(dev)
        dst_t = out_t[a // num_time:b // num_time] if want_device_t else None
        K.minmax_transform(chunk, scale, minv, out=chunk if want_host else None, out_t=dst_t,
                           T=num_time if want_device_t else 0)
        if want_host:
            flat[a:b] = chunk.cpu().numpy()
        del chunk
    return scaler, out_t


def _save_scaler(scaler, save_path):
    if save_path:
        d = os.path.dirname(save_path)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(save_path, "wb") as f:
            pickle.dump(scaler, f)


def _budget(dev):
    free, _ = torch.cuda.mem_get_info(dev)
    return int(free * 0.45)


def data_scaler(FOM_data_aug, FOM_data, num_time, num_node, directory, chunk_size=None, device=None,
                save_path="./model_save/scaler.pkl", chunk_bytes=2 << 30):
    """Drop-in for data_preprocess.data_scaler (data_preprocess.py:65-165): same arguments (FOM_data, directory and
    chunk_size are accepted and, like in the reference's arithmetic, unused), same return value
    `(new_x_train, DATA_shape, scaler)`; `FOM_data_aug` [P, T, N] is scaled in place when it is C-contiguous (the
    reference's `reshape(-1, num_node)` view has the same condition) and the scaler is pickled to
    ./model_save/scaler.pkl."""
    dev = _device(device)
    x = np.asarray(FOM_data_aug)
    flat = x.reshape(-1, num_node)
    scaler, _ = _scan(flat, int(num_time), dev, want_host=True, want_device_t=False, chunk_bytes=chunk_bytes,
                      resident_bytes=_budget(dev))
    new_x_train = flat.reshape(x.shape)
    _save_scaler(scaler, save_path)
    return new_x_train, new_x_train.shape[1:], scaler


def data_scaler_to_device(FOM_data_aug, num_time, num_node, device=None, save_path="./model_save/scaler.pkl",
                          chunk_bytes=2 << 30, scale_host_copy=False):
    """B200-first variant: returns `(x_train, DATA_shape, scaler)` with `x_train` the float32 `[P, N, T]` tensor in
    HBM - the result of data_scaler() followed by SimulGen-VAE.py:282-283, bit for bit - ready for
    `create_augmented_dataloaders(x_train, ..., load_all=True)`.  The host array is left untouched unless
    `scale_host_copy` is set."""
    dev = _device(device)
    x = np.asarray(FOM_data_aug)
    flat = x.reshape(-1, num_node)
    scaler, out_t = _scan(flat, int(num_time), dev, want_host=bool(scale_host_copy), want_device_t=True,
                          chunk_bytes=max(chunk_bytes, num_time * num_node * x.dtype.itemsize),
                          resident_bytes=_budget(dev))
    _save_scaler(scaler, save_path)
    return out_t, x.shape[1:], scaler
