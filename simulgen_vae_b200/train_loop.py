"""Drop-in for the reference's training driver `train()` (modules/train.py:49-245) on top of `Trainer`.

Same signature, side effects (./checkpoints/SimulGen-VAE.pth state dict, ./model_save/SimulGen-VAE whole-module
pickle, the per-epoch log line) and return value `(loss_print, recon_print, kl_print, loss_val_print)` as the
reference, same schedules (WarmupKLLoss beta per epoch, train.py:18-41,75-81; AdamW with torch defaults + weight decay
0.01, train.py:92; CosineAnnealingWarmRestarts(T_0=epochs//4, T_mult=2, eta_min=LR*1e-4) stepped per epoch,
train.py:94-96,237; validation every 20th and on the last epoch, train.py:181), but the batch loop is `Trainer.step`:
forward + backward through the engine, the fused spectral-norm-gradient + AdamW + grad-norm pass, NCCL all-reduce when
`torch.distributed` is initialised - and no host synchronisation inside an epoch (the reference reads 4 scalars and
126 per-parameter norms with .item() every step, train.py:156-174): the running sums live on the device and are read
once per epoch."""
from __future__ import annotations

import logging
import os
import time

import numpy as np
import torch

from . import engine
from .trainer import Trainer, warmup_beta


class WarmupKLLoss:
    """Same interface as the reference class (train.py:18-41): get_loss(step, losses) -> [beta, sum(losses)]."""

    def __init__(self, epoch, init_beta, start_warmup, end_warmup, beta_target):
        self.epoch, self.init_beta, self.beta_target = epoch, init_beta, beta_target
        self.start_warmup, self.end_warmup = start_warmup, end_warmup

    def beta(self, step):
        if step < self.start_warmup:
            return self.init_beta
        if step < self.end_warmup:
            return (step - self.start_warmup) * (self.beta_target - self.init_beta) / (self.end_warmup - self.start_warmup) \
                + self.init_beta
        return self.beta_target

    def get_loss(self, step, losses):
        total = 0
        for l in losses:
            total = total + l
        return [self.beta(step), total]


def print_gpu_mem_checkpoint(msg, debug_mode=0):
    if debug_mode == 1 and torch.cuda.is_available():
        print("[GPU MEM] %s: Allocated=%.2fMB, Max Allocated=%.2fMB" % (
            msg, torch.cuda.memory_allocated() / 1024 ** 2, torch.cuda.max_memory_allocated() / 1024 ** 2))
        torch.cuda.reset_peak_memory_stats()


def _lr_schedule(LR, epochs):
    """The reference's scheduler, stepped on a stand-in optimiser so that the learning rates are torch's own."""
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=LR)
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=epochs // 4, T_mult=2, eta_min=LR * 0.0001)
    return opt, sched


def train(epochs, batch_size, train_dataloader, val_dataloader, LR, num_filter_enc, num_filter_dec, num_node, latent_dim,
          hierarchical_dim, num_time, alpha, lossfun, small, load_all, debug_mode=0, device=None, model=None):
    """See the module docstring.  Extra keyword arguments (not in the reference): `device` (default cuda:0 / the local
    rank's GPU) and `model` (continue training an existing engine VAE instead of building a fresh one)."""
    from modules.VAE_network import VAE                      # the overlay's module classes
    from modules.common import add_sn, initialize_weights_He
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    os.makedirs("checkpoints", exist_ok=True)
    os.makedirs("output", exist_ok=True)
    os.makedirs("model_save", exist_ok=True)
    dist = torch.distributed
    ddp = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if ddp else 0
    if device is None:
        device = "cuda:%d" % (int(os.environ.get("LOCAL_RANK", "0")) if ddp else 0)
    device = torch.device(device)
    if device.type == "cuda":
        torch.cuda.set_device(device)
    if model is None:
        model = VAE(latent_dim, hierarchical_dim, num_filter_enc, num_filter_dec, num_node, num_time, lossfun=lossfun,
                    batch_size=batch_size, small=small, use_checkpointing=False)
        model.apply(initialize_weights_He)
        model.apply(add_sn)
    n_params = sum(p.numel() for p in model.parameters())
    if rank == 0:
        print("SimulGen-VAE on the B200 engine (%s, precision %s): %.1f M parameters, %d ranks" %
              (device, engine.get_precision(), n_params / 1e6, dist.get_world_size() if ddp else 1))
    warmup = WarmupKLLoss(epochs, 1e-4, int(epochs * 0.3), int(epochs * 0.8), 1)
    model.to(device)
    model.train(True)
    trainer = Trainer(model, lr=LR, alpha=alpha)
    sched_opt, scheduler = _lr_schedule(LR, epochs)

    loss_print, loss_val_print = np.zeros(epochs), np.zeros(epochs)
    recon_print, kl_print = np.zeros(epochs), np.zeros(epochs)
    recon_loss_MSE_print, recon_loss_val_print = np.zeros(epochs), np.zeros(epochs)
    for epoch in range(epochs):
        t0 = time.time()
        model.train(True)
        beta = warmup.beta(epoch)
        assert beta == warmup_beta(epoch, epochs)
        trainer.lr = sched_opt.param_groups[0]["lr"]
        acc = torch.zeros(5, dtype=torch.float64, device=device)         # loss, recon*alpha, kl*beta, mse*alpha, |grad|
        n_batches = 0
        for image in train_dataloader:
            if not load_all or image.device != device:
                image = image.to(device, non_blocking=True)
            # data parallel: the noise stream is keyed on the GLOBAL sample index (rank-major within a step)
            loss, recon, kl_sum, mse = trainer.step(image, beta=beta, sample_offset=rank * image.shape[0])
            acc += torch.stack([loss.double().reshape(()), recon.double().reshape(()) * alpha, kl_sum.double().reshape(()) * beta,
                                mse.double().reshape(()) * alpha, trainer.gnorm_sq.sqrt().reshape(())])
            n_batches += 1
        if ddp:                                                           # logged scalars are means over the GLOBAL batches
            dist.all_reduce(acc)
            acc /= dist.get_world_size()
        sums = acc.tolist()                                               # the epoch's only device -> host read
        if epoch % 20 == 0 or epoch == epochs - 1:
            trainer.sync()
            model.eval()
            vacc = torch.zeros(2, dtype=torch.float64, device=device)
            n_val = 0
            with torch.no_grad():
                for image in val_dataloader:
                    if not load_all or image.device != device:
                        image = image.to(device, non_blocking=True)
                    _, recon, kls, _ = model(image)
                    _, kl_sum = warmup.get_loss(epoch, kls)
                    r = recon.double().reshape(()) * alpha
                    vacc += torch.stack([r + kl_sum.double().reshape(()) * beta, r])
                    n_val += 1
            if ddp:
                dist.all_reduce(vacc)
                vacc /= dist.get_world_size()
            v = vacc.tolist()
            loss_val_print[epoch] = v[0] / max(n_val, 1)
            recon_loss_val_print[epoch] = v[1] / max(n_val, 1)
            model.train(True)
        elif epoch > 0:
            loss_val_print[epoch] = loss_val_print[epoch - 1]
            recon_loss_val_print[epoch] = recon_loss_val_print[epoch - 1]
        nb = max(n_batches, 1)
        loss_print[epoch] = sums[0] / nb
        recon_print[epoch] = sums[1] / nb
        kl_print[epoch] = sums[2] / beta / nb
        recon_loss_MSE_print[epoch] = sums[3] / nb
        current_lr = sched_opt.param_groups[0]["lr"]
        scheduler.step()
        dt = time.time() - t0
        if rank == 0:
            logging.info("\r[Epoch %d/%d] Loss: %.4E   val_loss: %.2E   Recon:%.4E   Recon_val:%.4E   KL:%.4E   Beta:%.4E   "
                         "AvgGrad:%.4E   Time: %.2fs   ETA: %.2fh    LR: %.2E" % (
                             epoch + 1, epochs, loss_print[epoch], loss_val_print[epoch], recon_print[epoch],
                             recon_loss_val_print[epoch], kl_print[epoch], beta, sums[4] / nb, dt,
                             (epochs - epoch) * dt / 3600, current_lr))
    trainer.sync()
    if rank == 0:
        torch.save(model.state_dict(), "checkpoints/SimulGen-VAE.pth")
        torch.save(model, "model_save/SimulGen-VAE")
    if device.type == "cuda":
        torch.cuda.empty_cache()
    return loss_print, recon_print, kl_print, loss_val_print
