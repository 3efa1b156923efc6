"""SURVEY §8f N3 - the inference / latent-export sweep of the reference, batched.

Reference: `evaluate_vae_reconstruction()` (modules/utils.py:428-561) walks a DataLoader - for the whole-dataset export
one with batch_size=1 (SimulGen-VAE.py:326-337) - and per parameter set runs `mu, log_var, xs = VAE.encoder(x)`,
`recon_iter` times `z = reparameterize(mu, exp(0.5 log_var))`, `gen_x, _ = VAE.decoder(z, xs, mode='fix')`,
`MSELoss(gen_x, x)`, keeps the best draw, and copies everything to the host sample by sample; SimulGen-VAE.py:339-344
then writes `model_save/latent_vectors.npy` [P, latent_dim_end], `model_save/xs.npy` [P, levels-1, latent_dim] and
`SimulGen-VAE_L2_loss.txt` ('%e').

Every op of the model is per sample (GroupNorm, not BatchNorm), so the sweep is regrouped into batches of `batch_size`
samples - one encoder and `recon_iter` decoder passes per batch through the engine's kernels, per-sample MSE on the
device, one device->host copy per batch - with the reference's signature, return values and file formats.
Differences, both deliberate: (1) a DataLoader with batch_size > 1 has ALL its samples recorded at consecutive rows
(the reference records only sample 0 of each batch); (2) the reparameterisation noise comes from the engine's
counter-based Philox stream keyed on the sample's position in the sweep, so results do not depend on `batch_size`
(the reference's come from torch's global generator in visiting order)."""
from __future__ import annotations

import os

import numpy as np
import torch


def _reparameterize():
    from modules.decoder import reparameterize        # the overlay's (engine Philox on CUDA)
    return reparameterize


def _regroup(dataloader, batch_size):
    """Yield tensors of up to batch_size samples, concatenating the loader's (possibly batch-1) batches in order."""
    pend, n = [], 0
    for image in dataloader:
        pend.append(image)
        n += image.shape[0]
        if n >= batch_size:
            yield torch.cat(pend) if len(pend) > 1 else pend[0]
            pend, n = [], 0
    if pend:
        yield torch.cat(pend) if len(pend) > 1 else pend[0]


def evaluate_vae_reconstruction(VAE, dataloader, device, num_param, num_filter_enc, latent_dim, latent_dim_end,
                                recon_iter=1, dataset_name="Dataset", save_images=True, batch_size=32,
                                keep_reconstructed=True, verbose=True):
    """Drop-in for utils.evaluate_vae_reconstruction (utils.py:428-561): returns
    (latent_vectors [num_param, latent_dim_end], hierarchical_latent_vectors [num_param, levels-1, latent_dim],
    reconstruction_loss [num_param], reconstructed [num_param, N, T], loss_total).  `batch_size`, `keep_reconstructed`
    (skip the [P, N, T] float64 host copy of every reconstruction) and `verbose` are extensions."""
    from . import engine
    reparameterize = _reparameterize()
    device = torch.device(device)
    save_dir = None
    if save_images:
        save_dir = "checkpoints/%s" % dataset_name.replace(" ", "_").replace("(", "").replace(")", "").lower()
        os.makedirs(save_dir, exist_ok=True)
    n_levels = len(num_filter_enc) - 1
    if verbose:
        print("Evaluating %s..." % dataset_name)
    # the sweep uses its own position-keyed noise stream; the training stream (seed, draw counter, sample offset) of
    # this thread is put back afterwards so that training which continues after an evaluation does not replay draws
    st = engine._rng_state()
    saved_rng = (st.seed, st.counter, st.sample0)
    try:
        return _sweep(VAE, dataloader, device, num_param, n_levels, latent_dim, latent_dim_end, recon_iter, dataset_name,
                      save_dir, batch_size, keep_reconstructed, verbose, reparameterize)
    finally:
        st.seed, st.counter, st.sample0 = saved_rng


def _sweep(VAE, dataloader, device, num_param, n_levels, latent_dim, latent_dim_end, recon_iter, dataset_name, save_dir,
           batch_size, keep_reconstructed, verbose, reparameterize):
    from . import engine
    latent_vectors = np.zeros([num_param, latent_dim_end])
    hierarchical = np.zeros([num_param, n_levels, latent_dim])
    reconstruction_loss = np.zeros([num_param])
    reconstructed = None
    loss_total = 0.0
    row = 0
    with torch.no_grad():
        for x in _regroup(dataloader, batch_size):
            x = x.to(device)
            Bb = x.shape[0]
            if reconstructed is None and keep_reconstructed:
                reconstructed = np.empty([num_param, x.shape[1], x.shape[2]])
            engine.set_sample_offset(row)                                   # noise keyed on the position in the sweep
            st = engine._rng_state()
            st.seed, st.counter = torch.initial_seed(), 0                   # ... and on the draw index within the sample
            mu, log_var, xs = VAE.encoder(x)
            best = torch.full((Bb,), 100.0, device=device, dtype=torch.float32)   # the reference's loss_save[:] = 100
            best_z = torch.zeros(Bb, mu.shape[1], device=device, dtype=mu.dtype)
            best_x = torch.zeros_like(x) if keep_reconstructed else None
            taken = torch.zeros(Bb, dtype=torch.bool, device=device)
            last = None
            for _ in range(recon_iter):
                std = torch.exp(0.5 * log_var)
                z = reparameterize(mu, std)
                gen_x, _ = VAE.decoder(z, xs, mode="fix")
                last = ((gen_x.float() - x.float()) ** 2).mean(dim=(1, 2))
                better = last < best
                best = torch.where(better, last, best)
                best_z = torch.where(better[:, None], z, best_z)
                if keep_reconstructed:
                    best_x = torch.where(better[:, None, None], gen_x, best_x)
                taken |= better
            n = min(Bb, num_param - row)
            if n > 0:
                tk = taken[:n].cpu().numpy()
                rows = np.arange(row, row + n)[tk]
                latent_vectors[rows] = best_z[:n].double().cpu().numpy()[tk]
                for k in range(min(len(xs), n_levels)):
                    hierarchical[rows, k, :] = xs[k][:n].double().cpu().numpy()[tk]
                reconstruction_loss[rows] = best[:n].double().cpu().numpy()[tk]
                if keep_reconstructed:
                    reconstructed[rows] = best_x[:n].double().cpu().numpy()[tk]
            last_host = last.double().cpu().numpy()
            loss_total += float(last_host.sum())
            if verbose:
                for j in range(Bb):
                    print("Parameter %d finished - MSE: %.4E" % (row + j + 1, last_host[j]))
            if save_dir is not None and row < 10 and keep_reconstructed:
                _save_plots(save_dir, x, reconstructed, row, min(n, 10 - row), last_host)
            row += Bb
    if verbose:
        print("\nTotal %s MSE loss: %.3e\n--------------------------------\n" % (dataset_name, loss_total / max(row, 1)))
    if reconstructed is None:
        reconstructed = np.empty([num_param, 0, 0])
    return latent_vectors, hierarchical, reconstruction_loss, reconstructed, loss_total


def _save_plots(save_dir, x, reconstructed, row0, count, losses):
    """Original vs reconstruction of the first channels of the first 10 samples (utils.py:519-545); skipped with a note
    when matplotlib is missing."""
    try:
        import matplotlib.pyplot as plt
    except Exception as e:  # pragma: no cover
        print("Warning: reconstruction images not saved (%s)" % e)
        return
    for j in range(count):
        original = x[j].float().cpu().numpy()
        recon = reconstructed[row0 + j]
        nch = min(3, original.shape[0])
        plt.figure(figsize=(12, 6))
        for ch in range(nch):
            plt.subplot(nch, 1, ch + 1)
            plt.plot(original[ch], label="Original", alpha=0.7)
            plt.plot(recon[ch], label="Reconstructed", alpha=0.7, linestyle="--")
            plt.title("Channel %d - Sample %d - MSE: %.4E" % (ch + 1, row0 + j + 1, losses[j]))
            plt.legend()
            plt.grid(True, alpha=0.3)
        plt.tight_layout()
        plt.savefig("%s/reconstruction_sample_%03d.png" % (save_dir, row0 + j + 1), dpi=300, bbox_inches="tight")
        plt.close()


def export_latents(VAE, x_data, device, num_filter_enc, latent_dim, latent_dim_end, recon_iter=1, batch_size=32,
                   out_dir="model_save", loss_file="./SimulGen-VAE_L2_loss.txt", save_images=False):
    """The whole-dataset export of SimulGen-VAE.py:326-344 in one call: x_data [P, N, T] (numpy or tensor, host or
    device) -> model_save/latent_vectors.npy, model_save/xs.npy, SimulGen-VAE_L2_loss.txt; returns the three arrays."""
    x = torch.as_tensor(x_data)
    P = x.shape[0]
    loader = (x[i:i + batch_size] for i in range(0, P, batch_size))
    lat, hier, rloss, _, _ = evaluate_vae_reconstruction(VAE, loader, device, P, num_filter_enc, latent_dim, latent_dim_end,
                                                         recon_iter, "Whole Dataset", save_images, batch_size,
                                                         keep_reconstructed=False, verbose=False)
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, "latent_vectors"), lat)
    np.save(os.path.join(out_dir, "xs"), hier)
    np.savetxt(loss_file, rloss, fmt="%e")
    return lat, hier, rloss
