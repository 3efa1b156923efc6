"""the batched inference / latent-export sweep (simulgen_vae_b200.export) against the reference's own
evaluate_vae_reconstruction (modules/utils.py:428-561) run with a batch-1 DataLoader on the same overlay model, with the
reparameterisation noise forced to zero on both sides; plus the export file formats of SimulGen-VAE.py:339-344.
CPU: kernels replaced by their torch models.  Skipped where the reference checkout is absent."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


def _purge_modules():
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]


@pytest.mark.parametrize("batch_size", [1, 4, 5])
def test_batched_export_matches_reference_sweep(batch_size, tmp_path, monkeypatch):
    ref_import._install_stubs()
    monkeypatch.chdir(tmp_path)
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)
    g = torch.Generator().manual_seed(1)
    data = torch.rand(9, cfg["num_node"], cfg["num_time"], generator=g) * 1.4 - 0.7
    monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: torch.zeros_like(t))
    _purge_modules()
    sg.install_overlay(train=True)
    sys.path.insert(2, ref_import.REFERENCE_ROOT)
    sg.set_precision("fp32")
    try:
        utils = importlib.import_module("modules.utils")
        assert utils.__file__.startswith(sg.OVERLAY_TRAIN_DIR)
        assert utils.reference_evaluate_vae_reconstruction.__code__.co_filename.startswith(ref_import.REFERENCE_ROOT)
        from simulgen_vae_b200 import export
        assert utils.evaluate_vae_reconstruction is export.evaluate_vae_reconstruction
        assert callable(utils.parse_condition_file) and callable(utils.get_optimal_workers)      # re-exported names
        from modules.VAE_network import VAE
        from modules.common import add_sn, initialize_weights_He
        torch.manual_seed(0)
        m = VAE(cfg["latent_dim"], cfg["hierarchical_dim"], cfg["enc"], cfg["enc"][::-1], cfg["num_node"], cfg["num_time"],
                lossfun="MSE", batch_size=4, small=True)
        m.apply(initialize_weights_He)
        m.apply(add_sn)
        m.eval()
        args = ("cpu", 9, cfg["enc"], cfg["hierarchical_dim"], cfg["latent_dim"])
        with emu.install():
            loader1 = torch.utils.data.DataLoader(utils.Dataset(data.numpy(), False), batch_size=1, shuffle=False)
            ref = utils.reference_evaluate_vae_reconstruction(m, loader1, *args, recon_iter=2, dataset_name="ref", save_images=False)
            loader1 = torch.utils.data.DataLoader(utils.Dataset(data.numpy(), False), batch_size=1, shuffle=False)
            ours = export.evaluate_vae_reconstruction(m, loader1, *args, recon_iter=2, dataset_name="ours", save_images=False,
                                                      batch_size=batch_size, verbose=False)
            lat, hier, rloss = export.export_latents(m, data.numpy(), "cpu", cfg["enc"], cfg["hierarchical_dim"],
                                                     cfg["latent_dim"], recon_iter=1, batch_size=batch_size)
        names = ("latent_vectors", "hierarchical_latent_vectors", "reconstruction_loss", "reconstructed")
        for name, a, b in zip(names, ref[:4], ours[:4]):
            assert a.shape == b.shape and a.dtype == b.dtype, name
            assert np.allclose(a, b, rtol=1e-4, atol=1e-6), (name, np.abs(a - b).max())
        assert abs(float(ref[4]) - float(ours[4])) < 1e-4 * abs(float(ref[4]))
        # file formats of SimulGen-VAE.py:339-344
        assert np.load("model_save/latent_vectors.npy").shape == (9, cfg["latent_dim"])
        assert np.load("model_save/xs.npy").shape == (9, len(cfg["enc"]) - 1, cfg["hierarchical_dim"])
        txt = np.loadtxt("SimulGen-VAE_L2_loss.txt")
        assert txt.shape == (9,) and np.allclose(txt, rloss, rtol=1e-6)
        assert np.allclose(lat, ref[0], rtol=1e-4, atol=1e-6) and np.allclose(hier, ref[1], rtol=1e-4, atol=1e-6)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
        sg.install_overlay(train=False)
        if ref_import.REFERENCE_ROOT in sys.path:
            sys.path.remove(ref_import.REFERENCE_ROOT)
        _purge_modules()
