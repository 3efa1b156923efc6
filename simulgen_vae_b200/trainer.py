"""Training-step driver for the engine: the reference's step semantics (train.py:139-168) without
its host synchronisations, plus data-parallel gradient all-reduce over NCCL.

  step(x):  zero grads -> VAE.forward -> alpha*recon + beta*sum(kl) -> backward
            -> (DP) bucketed all-reduce of the gradients, overlapped with the rest of backward
            -> fused AdamW + global grad-norm (sg_adamw_step), AdamW defaults of torch (train.py:92).

The reference computes the gradient norm with one `.item()` per parameter (126 blocking D2H copies
per step, train.py:156-161) and four more for logging (train.py:171-174); here all scalars stay on
the device and are fetched only when the caller asks (`Trainer.scalars()`).
"""
from __future__ import annotations

import torch

from . import engine
from . import kernels as K


def warmup_beta(epoch: int, epochs: int, init_beta: float = 1e-4, beta_target: float = 1.0) -> float:
    """WarmupKLLoss schedule of the reference (train.py:18-41,75-81): a function of the epoch."""
    start, end = int(epochs * 0.3), int(epochs * 0.8)
    if epoch < start:
        return init_beta
    if epoch < end:
        return (epoch - start) * (beta_target - init_beta) / (end - start) + init_beta
    return beta_target


class Trainer:
    def __init__(self, model, lr=1e-3, alpha=1.0e6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                 process_group=None, bucket_mb=64):
        self.model = model
        self.lr, self.alpha, self.betas, self.eps, self.wd = lr, alpha, betas, eps, weight_decay
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.m = {}
        self.v = {}
        self.step_count = 0
        dev = self.params[0].device
        self.gnorm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.bucket_bytes = bucket_mb * (1 << 20)
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == "cuda") else None
        self._last = None

    # -- data parallel --------------------------------------------------------------------------
    def _allreduce_grads(self):
        """Bucketed sum all-reduce of the local gradients (mean is folded into AdamW's grad_scale).
        Gradients are taken in reverse parameter order = the order backward produced them."""
        if self.world == 1:
            return
        dist = torch.distributed
        grads = [p.grad for p in reversed(self.params) if p.grad is not None]
        bucket, size, handles = [], 0, []
        for g in grads:
            bucket.append(g)
            size += g.numel() * 4
            if size >= self.bucket_bytes:
                handles.append(self._launch_bucket(bucket))
                bucket, size = [], 0
        if bucket:
            handles.append(self._launch_bucket(bucket))
        for flat, bucket, work in handles:
            work.wait()
            off = 0
            for g in bucket:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n

    def _launch_bucket(self, bucket):
        dist = torch.distributed
        if len(bucket) == 1:
            flat = bucket[0].view(-1)
        else:
            flat = torch.cat([g.reshape(-1) for g in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        return flat, ([] if len(bucket) == 1 else bucket), work

    # -- one optimisation step -------------------------------------------------------------------
    def step(self, x, beta=1e-4, sample_offset=0):
        model = self.model
        for p in self.params:
            p.grad = None
        engine.set_sample_offset(sample_offset)
        x_hat, recon, kls, mse = model(x)
        kl_sum = kls[0]
        for k in kls[1:]:
            kl_sum = kl_sum + k
        loss = recon * self.alpha + kl_sum * beta
        loss.backward()
        self._allreduce_grads()
        self.step_count += 1
        self.gnorm_sq.zero_()
        b1, b2 = self.betas
        scale = 1.0 / self.world
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            st = self.m.get(p)
            if st is None:
                self.m[p] = torch.zeros_like(p)
                self.v[p] = torch.zeros_like(p)
            K.adamw_step(p.data, g, self.m[p], self.v[p], self.lr, b1, b2, self.eps, self.wd, self.step_count, scale,
                         self.gnorm_sq)
        self._last = (loss.detach(), recon.detach(), kl_sum.detach(), mse.detach())
        return self._last

    def scalars(self):
        """(loss, recon, kl_sum, mse, grad_norm) as Python floats - the only host synchronisation."""
        loss, recon, kl, mse = self._last
        return float(loss), float(recon), float(kl), float(mse), float(self.gnorm_sq.sqrt())
