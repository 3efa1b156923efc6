"""-m gpu: the drop-in claim on REAL kernels.  The UNMODIFIED reference drivers - modules/train.py
::train() and modules/utils.py::evaluate_vae_reconstruction(), loaded from the staged byte-for-byte copy oracle/_ref on
the GPU box (oracle/make_ref.sh) or from /root/reference - run on top of the engine's overlay modules with the CUDA
extension doing the arithmetic (no kernel emulator here)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import simulgen_vae_b200 as sg
from oracle import ref_import

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="reference checkout (or its staged copy oracle/_ref) not present")]

CFG = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)


def _purge_modules():
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]


# static fields (Dim2 = 1): batch 8 so that the 16-bit modes take the compact [C][B] path of the two N-channel layers
CFG_STATIC = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=200, num_time=1)


def _data(cfg=CFG, samples=16):
    g = torch.Generator().manual_seed(0)
    return torch.rand(samples, cfg["num_node"], cfg["num_time"], generator=g) * 1.4 - 0.7


@pytest.mark.parametrize("precision,shape", [("fp32", "T20"), ("fp16", "T20"), ("fp16", "static")])
def test_unmodified_reference_train_runs_on_real_kernels(precision, shape, tmp_path, monkeypatch):
    """reference train(): model.apply(initialize_weights_He / add_sn), AdamW over model.parameters(), loss.backward()
    through the engine's autograd Functions, the per-parameter grad-norm loop, validation under no_grad, torch.save of the
    state dict and of the whole module - on cuda:0.  In fp32 mode the loss curve must equal the engine's own train() driver
    (Trainer: direct tapes + fused optimiser) on the same seeds."""
    ref_import._install_stubs()
    monkeypatch.chdir(tmp_path)
    os.makedirs("checkpoints")
    os.makedirs("model_save")
    CFG = CFG_STATIC if shape == "static" else globals()["CFG"]
    bs, n_train, n_all = (8, 16, 24) if shape == "static" else (4, 12, 16)
    data = _data(CFG, n_all).cuda()
    # alpha: the reference's train() has no loss scaling, so in fp16 mode alpha / numel must keep the gradients of the
    # 16-bit operands in range; the static toy field has 1600 elements per batch -> alpha = 1e4
    alpha = 10000 if shape == "static" else 1000000
    args = (5, bs, None, None, 1e-3, CFG["enc"], CFG["enc"][::-1], CFG["num_node"], CFG["latent_dim"], CFG["hierarchical_dim"],
            CFG["num_time"], alpha, "MSE", True, True)
    seen = {"static_fwd": 0}
    from simulgen_vae_b200 import kernels as K
    orig_sf = K.static_recon_fwd

    def counting_sf(*a, **k):
        seen["static_fwd"] += 1
        return orig_sf(*a, **k)
    monkeypatch.setattr(K, "static_recon_fwd", counting_sf)

    def run(use_engine_driver):
        _purge_modules()
        sg.install_overlay(train=use_engine_driver)
        sys.path.insert(2, ref_import.REFERENCE_ROOT)
        try:
            train_mod = importlib.import_module("modules.train")
            vae_mod = importlib.import_module("modules.VAE_network")
            assert vae_mod.__file__.startswith(sg.OVERLAY_DIR)
            assert train_mod.__file__.startswith(sg.OVERLAY_TRAIN_DIR if use_engine_driver else ref_import.REFERENCE_ROOT)
            from simulgen_vae_b200 import engine
            torch.manual_seed(3)
            engine._rng_state().seed = None
            a = list(args)
            a[2] = torch.utils.data.DataLoader(data[:n_train], batch_size=bs, shuffle=False)
            a[3] = torch.utils.data.DataLoader(data[n_train:], batch_size=bs, shuffle=False)
            curves = train_mod.train(*a)
            sd = torch.load("checkpoints/SimulGen-VAE.pth", weights_only=False)
            whole = torch.load("model_save/SimulGen-VAE", weights_only=False)
            assert type(whole).__module__ == "modules.VAE_network" and type(whole).__name__ == "VAE"
            return [np.asarray(c, dtype=np.float64) for c in curves], sd
        finally:
            sg.install_overlay(train=False)
            if ref_import.REFERENCE_ROOT in sys.path:
                sys.path.remove(ref_import.REFERENCE_ROOT)
            _purge_modules()

    sg.set_precision(precision)
    try:
        ref_curves, ref_sd = run(False)
        assert all(np.isfinite(c).all() for c in ref_curves)
        ref_model = ref_import.build_reference_vae(dict(CFG, batch=bs, small=True, lossfun="MSE"))
        assert list(ref_sd.keys()) == list(ref_model.state_dict().keys())
        ref_model.load_state_dict({k: v.cpu() for k, v in ref_sd.items()})       # engine checkpoint loads into the reference
        eng_curves, eng_sd = run(True)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    # static fields in a 16-bit mode: both drivers went through the compact head (train + validation forwards)
    assert (seen["static_fwd"] > 0) == (shape == "static"), seen
    tol = 2e-4 if precision == "fp32" else 2e-2
    for name, a, b in zip(("loss", "recon", "kl", "val_loss"), ref_curves, eng_curves):
        assert a.shape == b.shape == (5,)
        assert np.allclose(a, b, rtol=tol, atol=1e-6), (name, a, b)


def test_export_sweep_matches_reference_on_real_kernels(tmp_path, monkeypatch):
    """utils.evaluate_vae_reconstruction of the UNMODIFIED reference (batch-1 DataLoader) against the engine's batched sweep
    on the same overlay model on cuda:0, reparameterisation noise forced to zero on both sides (fp32 validation mode)."""
    ref_import._install_stubs()
    monkeypatch.chdir(tmp_path)
    data = _data()[:9]
    monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: torch.zeros_like(t))
    _purge_modules()
    sg.install_overlay(train=True)
    sys.path.insert(2, ref_import.REFERENCE_ROOT)
    sg.set_precision("fp32")
    try:
        utils = importlib.import_module("modules.utils")
        from simulgen_vae_b200 import export
        from modules.VAE_network import VAE
        from modules.common import add_sn, initialize_weights_He
        torch.manual_seed(0)
        m = VAE(CFG["latent_dim"], CFG["hierarchical_dim"], CFG["enc"], CFG["enc"][::-1], CFG["num_node"], CFG["num_time"],
                lossfun="MSE", batch_size=4, small=True)
        m.apply(initialize_weights_He)
        m.apply(add_sn)
        m.cuda().eval()
        args = ("cuda", 9, CFG["enc"], CFG["hierarchical_dim"], CFG["latent_dim"])
        loader1 = torch.utils.data.DataLoader(utils.Dataset(data.numpy(), False), batch_size=1, shuffle=False)
        ref = utils.reference_evaluate_vae_reconstruction(m, loader1, *args, recon_iter=2, dataset_name="ref", save_images=False)
        loader1 = torch.utils.data.DataLoader(utils.Dataset(data.numpy(), False), batch_size=1, shuffle=False)
        ours = export.evaluate_vae_reconstruction(m, loader1, *args, recon_iter=2, dataset_name="ours", save_images=False,
                                                  batch_size=4, verbose=False)
        for name, a, b in zip(("latent_vectors", "hierarchical_latent_vectors", "reconstruction_loss", "reconstructed"), ref[:4], ours[:4]):
            assert a.shape == b.shape, name
            assert np.allclose(a, b, rtol=1e-4, atol=1e-6), (name, np.abs(a - b).max())
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
        sg.install_overlay(train=False)
        if ref_import.REFERENCE_ROOT in sys.path:
            sys.path.remove(ref_import.REFERENCE_ROOT)
        _purge_modules()
