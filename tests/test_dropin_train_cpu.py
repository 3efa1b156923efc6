"""CPU test of the drop-in claim: the UNMODIFIED /root/reference/modules/train.py::train() runs on the
engine's overlay modules (namespace-package shadowing, `install_overlay()`), with the CUDA kernels replaced by their
torch models.  Checks that the reference's own driver - model.apply(initialize_weights_He / add_sn), AdamW over
model.parameters(), the per-parameter grad-norm loop, validation under no_grad, torch.save(state_dict) and
torch.save(model) - works against the engine and that the checkpoint layout equals the reference's.
Skipped where the reference checkout is absent (GPU box)."""
import importlib
import os
import sys

import pytest
import torch

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


def _purge_modules():
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]


def test_reference_train_runs_on_the_overlay(tmp_path, monkeypatch):
    ref_import._install_stubs()
    monkeypatch.chdir(tmp_path)
    os.makedirs("checkpoints")
    os.makedirs("model_save")
    _purge_modules()
    overlay = sg.install_overlay()
    sys.path.insert(1, ref_import.REFERENCE_ROOT)
    try:
        train_mod = importlib.import_module("modules.train")
        vae_mod = importlib.import_module("modules.VAE_network")
        assert vae_mod.__file__.startswith(overlay), "train.py must bind to the overlay's VAE"
        assert train_mod.__file__.startswith(ref_import.REFERENCE_ROOT), "train.py itself must be the reference's"
        assert train_mod.VAE is vae_mod.VAE
        cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)
        g = torch.Generator().manual_seed(0)
        data = torch.rand(16, cfg["num_node"], cfg["num_time"], generator=g) * 1.4 - 0.7
        train_dl = torch.utils.data.DataLoader(data[:12], batch_size=4, shuffle=False)
        val_dl = torch.utils.data.DataLoader(data[12:], batch_size=4, shuffle=False)
        sg.set_precision("fp32")
        monkeypatch.setattr(train_mod, "device", torch.device("cpu"), raising=False)
        with emu.install():
            out = train_mod.train(4, 4, train_dl, val_dl, 1e-3, cfg["enc"], cfg["enc"][::-1], cfg["num_node"],
                                  cfg["latent_dim"], cfg["hierarchical_dim"], cfg["num_time"], 1000000, "MSE", True, True)
        assert len(out) == 4 and all(len(c) == 4 for c in out)
        loss_curve = [float(v) for v in out[0]]
        assert all(v == v and v < 1e12 for v in loss_curve)            # finite
        sd = torch.load("checkpoints/SimulGen-VAE.pth", weights_only=False)
        # the checkpoint layout of the reference at this toy preset (weight_orig / weight_u / weight_v ...)
        ref = ref_import.build_reference_vae(dict(cfg, batch=4, small=True, lossfun="MSE"))
        assert list(sd.keys()) == list(ref.state_dict().keys())
        for k, v in ref.state_dict().items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
        ref.load_state_dict(sd)                                          # engine checkpoint loads into the reference
        whole = torch.load("model_save/SimulGen-VAE", weights_only=False)
        assert type(whole).__module__ == "modules.VAE_network" and type(whole).__name__ == "VAE"
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
        if ref_import.REFERENCE_ROOT in sys.path:
            sys.path.remove(ref_import.REFERENCE_ROOT)
        _purge_modules()


def test_engine_train_driver_matches_the_reference_driver(tmp_path, monkeypatch):
    """simulgen_vae_b200.train_loop.train (Trainer-based, one host read per epoch) against the reference's
    own train() on the same overlay modules, seeds and data: same curves, same final weights, same checkpoint files."""
    ref_import._install_stubs()
    monkeypatch.chdir(tmp_path)
    os.makedirs("checkpoints")
    os.makedirs("model_save")
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)
    g = torch.Generator().manual_seed(0)
    data = torch.rand(16, cfg["num_node"], cfg["num_time"], generator=g) * 1.4 - 0.7
    args = (5, 4, None, None, 1e-3, cfg["enc"], cfg["enc"][::-1], cfg["num_node"], cfg["latent_dim"], cfg["hierarchical_dim"],
            cfg["num_time"], 1000000, "MSE", True, True)

    def run(use_engine_driver):
        _purge_modules()
        sg.install_overlay(train=use_engine_driver)
        sys.path.insert(2, ref_import.REFERENCE_ROOT)
        try:
            train_mod = importlib.import_module("modules.train")
            assert train_mod.__file__.startswith(sg.OVERLAY_TRAIN_DIR if use_engine_driver else ref_import.REFERENCE_ROOT)
            from simulgen_vae_b200 import engine
            torch.manual_seed(3)
            engine._rng_state().seed = None
            a = list(args)
            a[2] = torch.utils.data.DataLoader(data[:12], batch_size=4, shuffle=False)
            a[3] = torch.utils.data.DataLoader(data[12:], batch_size=4, shuffle=False)
            kw = dict(device="cpu") if use_engine_driver else {}
            with emu.install():
                curves = train_mod.train(*a, **kw)
            sd = torch.load("checkpoints/SimulGen-VAE.pth", weights_only=False)
            whole = torch.load("model_save/SimulGen-VAE", weights_only=False)
            assert type(whole).__name__ == "VAE"
            return [np.asarray(c, dtype=np.float64) for c in curves], sd
        finally:
            sg.install_overlay(train=False)
            if ref_import.REFERENCE_ROOT in sys.path:
                sys.path.remove(ref_import.REFERENCE_ROOT)
            _purge_modules()

    import numpy as np
    sg.set_precision("fp32")
    try:
        ref_curves, ref_sd = run(False)
        eng_curves, eng_sd = run(True)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    for name, a, b in zip(("loss", "recon", "kl", "val_loss"), ref_curves, eng_curves):
        assert a.shape == b.shape == (5,)
        assert np.allclose(a, b, rtol=2e-4, atol=1e-6), (name, a, b)
    assert list(ref_sd.keys()) == list(eng_sd.keys())
    for k in ref_sd:
        a, b = ref_sd[k].double(), eng_sd[k].double()
        assert float((a - b).norm() / (a.norm() + 1e-30)) < 2e-3, k
