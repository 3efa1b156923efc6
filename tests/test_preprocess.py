"""Preprocessing scan against golden output of the UNMODIFIED reference data_scaler
(oracle/make_golden_scaler.py -> tests/golden/scaler_toy.npz) and against sklearn / NumPy directly.
CPU: the host driver (row sampling, chunking, partial-fit merging, scaler construction, pickle) with the two kernels
replaced by their torch models.  GPU (-m gpu): the real sg_minmax_fit / sg_minmax_transform - bit-exact."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

import kernel_emulator as emu
from conftest import GOLDEN_DIR, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_scaler import CASES, make_data  # noqa: E402  (the generator of the golden inputs; no reference needed)

ATTRS = ("data_min_", "data_max_", "data_range_", "scale_", "min_")


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "scaler_toy.npz"))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def _check_case(name, device, tmp_path, monkeypatch, chunk_bytes, resident):
    from simulgen_vae_b200 import preprocess as pp
    g = _golden()
    dtype, P, T, N = CASES[name]
    x = make_data(dtype, P, T, N, int(g[name + "_seed"]))
    if device == "cpu":
        monkeypatch.setattr(pp, "_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(pp, "_budget", lambda d: (1 << 40) if resident else 0)
    save = str(tmp_path / "model_save" / "scaler.pkl")
    work = x.copy()
    y, shape, scaler = pp.data_scaler(work, x, T, N, 1, chunk_size=64, device=device, save_path=save, chunk_bytes=chunk_bytes)
    assert y.base is work or y is work or np.shares_memory(y, work)            # scaled in place like the reference
    assert tuple(shape) == (T, N)
    assert _same(y, g[name + "_scaled"]), "scaled field differs from the reference's"
    for a in ATTRS:
        assert _same(getattr(scaler, a), g[name + "_" + a]), a
    assert scaler.n_samples_seen_ == int(g[name + "_n_samples_seen"])
    from sklearn.preprocessing import MinMaxScaler
    with open(save, "rb") as f:
        loaded = pickle.load(f)
    assert isinstance(loaded, MinMaxScaler) and _same(loaded.scale_, scaler.scale_)
    # inverse_transform of the pickled scaler works like the reference's evaluators expect (utils / evaluators use it)
    back = loaded.inverse_transform(y.reshape(-1, N)[:5].copy())
    assert np.allclose(back, x.reshape(-1, N)[:5], rtol=1e-4 if dtype == np.float32 else 1e-10, atol=1e-4, equal_nan=True)
    # B200-first entry: float32 [P, N, T] == np.float32(reference output transposed), host array untouched
    work2 = x.copy()
    xt, shape2, scaler2 = pp.data_scaler_to_device(work2, T, N, device=device, save_path=None, chunk_bytes=chunk_bytes)
    assert _same(work2, x) and tuple(shape2) == (T, N)
    want = np.float32(g[name + "_scaled"].transpose((0, 2, 1)))
    assert xt.dtype == torch.float32 and tuple(xt.shape) == (P, N, T)
    assert _same(xt.cpu().numpy(), np.ascontiguousarray(want))
    assert _same(scaler2.min_, scaler.min_)


@pytest.mark.parametrize("name", ["f64", "f32"])
@pytest.mark.parametrize("chunk_bytes,resident", [(1 << 30, True), (6000, False), (20000, True)])
def test_data_scaler_host_logic_matches_reference(name, chunk_bytes, resident, tmp_path, monkeypatch):
    with emu.install():
        _check_case(name, "cpu", tmp_path, monkeypatch, chunk_bytes, resident)


def test_reference_row_sampling():
    from simulgen_vae_b200 import preprocess as pp
    for total in (280, 1200, 96800, 600000):
        idx = pp.reference_sample_rows(total)
        max_samples = min(50000, total // 10)
        if max_samples < 1000:
            max_samples = min(1000, total)
        np.random.seed(42)
        want = np.random.choice(total, max_samples, replace=False) if total > max_samples else np.arange(total)
        assert np.array_equal(idx, want)


def test_overlay_module_keeps_the_reference_names(monkeypatch):
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference checkout not present")
    from oracle import ref_import
    import simulgen_vae_b200 as sg
    ref_import._install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    for k in saved:
        del sys.modules[k]
    monkeypatch.syspath_prepend(ref_root)
    sg.install_overlay()
    try:
        import importlib
        m = importlib.import_module("modules.data_preprocess")
        from simulgen_vae_b200 import preprocess as pp
        assert m.data_scaler is pp.data_scaler
        assert m.reference_data_scaler.__code__.co_filename.startswith(ref_root)
        for name in ("reduce_dataset", "latent_conditioner_scaler", "get_memory_usage"):
            assert callable(getattr(m, name)), name
    finally:
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if sg.OVERLAY_DIR in sys.path:
            sys.path.remove(sg.OVERLAY_DIR)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f64", "f32"])
@pytest.mark.parametrize("chunk_bytes,resident", [(1 << 30, True), (6000, False)])
def test_data_scaler_gpu_matches_reference(name, chunk_bytes, resident, tmp_path, monkeypatch):
    _check_case(name, "cuda", tmp_path, monkeypatch, chunk_bytes, resident)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("R,N,T", [(600, 1000, 200), (37, 129, 37), (64, 2, 8), (5, 4099, 1), (0, 16, 1)])
def test_minmax_kernels_bit_exact_vs_sklearn(dtype, R, N, T):
    from sklearn.preprocessing import MinMaxScaler
    from simulgen_vae_b200 import kernels as K
    rng = np.random.default_rng(R * 7 + N)
    x = (rng.normal(size=(R, N)) * rng.uniform(0.01, 100, size=N) + rng.uniform(-3, 3, size=N)).astype(dtype)
    if R > 2:
        x[:, 0] = 0.5                        # constant feature
        x[1, N - 1] = np.nan
    dev = torch.device("cuda")
    xd = torch.from_numpy(x).to(dev)
    td = xd.dtype
    mn, mx = torch.empty(N, dtype=td, device=dev), torch.empty(N, dtype=td, device=dev)
    if R == 0:
        K.minmax_fit(xd, None, mn, mx)
        torch.cuda.synchronize()
        assert torch.isinf(mn).all() and torch.isinf(mx).all()
        return
    # all rows
    K.minmax_fit(xd, None, mn, mx)
    with np.errstate(all="ignore"):
        assert _same(mn.cpu().numpy(), np.nanmin(x, axis=0)) and _same(mx.cpu().numpy(), np.nanmax(x, axis=0))
    # a row subset in two merged pieces == fit on the gathered subset
    sel = np.sort(rng.choice(R, max(1, R // 3), replace=False))
    h = len(sel) // 2
    K.minmax_fit(xd, torch.from_numpy(sel[:h]).to(dev), mn, mx)
    K.minmax_fit(xd, torch.from_numpy(sel[h:]).to(dev), mn, mx, merge=h > 0)
    sc = MinMaxScaler(feature_range=(-0.7, 0.7)).fit(x[sel])
    assert _same(mn.cpu().numpy(), sc.data_min_) and _same(mx.cpu().numpy(), sc.data_max_)
    # transform: out of place, in place, and the float32 [P, N, T] layout
    want = sc.transform(x.copy())
    scale, minv = torch.from_numpy(sc.scale_).to(dev), torch.from_numpy(sc.min_).to(dev)
    out = torch.empty_like(xd)
    K.minmax_transform(xd, scale, minv, out=out)
    assert _same(out.cpu().numpy(), want)
    if R % T == 0:
        out_t = torch.empty(R // T, N, T, dtype=torch.float32, device=dev)
        inpl = xd.clone()
        K.minmax_transform(inpl, scale, minv, out=inpl, out_t=out_t, T=T)
        assert _same(inpl.cpu().numpy(), want)
        assert _same(out_t.cpu().numpy(), np.ascontiguousarray(np.float32(want.reshape(R // T, T, N).transpose(0, 2, 1))))
        only_t = torch.empty_like(out_t)
        K.minmax_transform(xd, scale, minv, out=None, out_t=only_t, T=T)
        assert _same(only_t.cpu().numpy(), out_t.cpu().numpy())
