"""Batch assembly + on-the-fly augmentation as ONE CUDA kernel per batch.

Replaces `modules/augmentation.py` of the reference for a dataset that is resident on the GPU
(`load_all=True`, SimulGen-VAE.py:290-295): there, every sample of a batch goes through Python
(`AugmentedDataset.__getitem__`, augmentation.py:43-84: optional Gaussian noise sigma=0.05, optional amplitude
scaling U[0.9,1.1], optional mixup with another random sample, lambda ~ Beta(0.2,0.2) clipped to [0.1,0.9]), each
step a separate full pass over a 76 MB sample, followed by `default_collate` (one more pass).  Here the host
only draws the per-sample DECISIONS - with Python's `random` / `numpy.random` in exactly the order the reference
draws them, so the same seeds give the same decisions - and `sg_assemble_batch` gathers, perturbs, mixes and
writes the batch in one pass; it can also emit the bf16 operand of the first encoder conv (fusing sg_pack_input).

The noise VALUES come from the kernel's counter-based Philox generator (the reference calls torch.randn_like per
sample); tests inject a noise tensor on both sides.  The order in which samples are visited is produced by a real
torch DataLoader over the index set, so the torch generator is consumed exactly as by the reference's DataLoader.
"""
from __future__ import annotations

import random

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Subset

from . import kernels as K

# augmentation.py:26-38 (the `augmentation_config` argument of the reference is ignored: these are always used)
DEFAULTS = dict(noise_prob=0.5, noise_level=0.05, scaling_prob=0.5, scaling_range=(0.9, 1.1), shift_prob=0.0,
                mixup_prob=0.5, mixup_alpha=0.2, cutout_prob=0.0)


class _IndexDataset(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return torch.as_tensor(i)


def draw_decisions(indices, dataset_len, cfg=DEFAULTS):
    """Per-sample augmentation decisions for one batch, consuming `random` and `numpy.random` exactly like
    AugmentedDataset._apply_augmentations (augmentation.py:57-84) called once per sample in batch order.
    Returns (noise_level[B], scale[B], other[B] (-1 = no mixup), lam[B], 1 - lam[B]); the scalars are formed in
    double precision (Python floats in the reference) and rounded to float32 once, as ATen does for a Python scalar."""
    B = len(indices)
    noise = np.zeros(B, np.float32)
    scale = np.ones(B, np.float32)
    other = np.full(B, -1, np.int64)
    lam = np.ones(B, np.float32)
    one_minus = np.zeros(B, np.float32)
    for i, index in enumerate(indices):
        index = int(index)
        if random.random() < cfg["noise_prob"]:
            noise[i] = cfg["noise_level"]
        if random.random() < cfg["scaling_prob"]:
            lo, hi = cfg["scaling_range"]
            scale[i] = lo + random.random() * (hi - lo)
        random.random()                                   # shift_prob = 0: the draw is still consumed
        if random.random() < cfg["mixup_prob"] and dataset_len > 1:
            o = random.randint(0, dataset_len - 1)
            while o == index:
                o = random.randint(0, dataset_len - 1)
            other[i] = o
            lm = max(0.1, min(float(np.random.beta(cfg["mixup_alpha"], cfg["mixup_alpha"])), 0.9))
            lam[i] = lm
            one_minus[i] = 1 - lm
        random.random()                                   # cutout_prob = 0: the draw is still consumed
    return noise, scale, other, lam, one_minus


class B200AugmentedLoader:
    """Iterable with the DataLoader surface train.py uses (`for image in loader`, `len(loader)`): yields fp32
    [B, N, T] CUDA batches assembled by sg_assemble_batch from the GPU-resident dataset."""

    def __init__(self, data, indices, batch_size, shuffle, augment, seed=0):
        self.data = data                                  # [P, N, T] fp32 CUDA, contiguous
        self.indices = torch.as_tensor(indices, dtype=torch.int64)
        self.batch_size, self.shuffle, self.augment = batch_size, shuffle, augment
        self.dataset_len = data.shape[0]
        self._index_loader = DataLoader(Subset(_IndexDataset(self.dataset_len), self.indices.tolist()),
                                        batch_size=batch_size, shuffle=shuffle, num_workers=0)
        self.seed, self.draws = seed, 0
        self.injected_noise = None                        # tests: callable(batch_position, shape) -> noise tensor
        self.last_decisions = None
        self.emit_operand = False                         # also produce the packed bf16 conv operand
        self.last_operand = None

    def __len__(self):
        return len(self._index_loader)

    def __iter__(self):
        dev = self.data.device
        for idx in self._index_loader:
            idx = idx.reshape(-1)
            B = idx.numel()
            if self.augment:
                noise, scale, other, lam, om = draw_decisions(idx.tolist(), self.dataset_len)
            else:
                noise, scale = np.zeros(B, np.float32), np.ones(B, np.float32)
                other, lam, om = np.full(B, -1, np.int64), np.ones(B, np.float32), np.zeros(B, np.float32)
            self.last_decisions = dict(index=idx.clone(), noise=noise, scale=scale, other=other, lam=lam)
            table = torch.from_numpy(np.stack([noise, scale, lam, om]).astype(np.float32)).to(dev, non_blocking=True)
            ids = torch.stack([idx, torch.from_numpy(other)]).to(torch.int32).to(dev, non_blocking=True)
            inj = None
            if self.injected_noise is not None and (noise > 0).any():
                inj = self.injected_noise(noise, tuple(self.data.shape[1:])).to(dev)
            out = torch.empty((B,) + tuple(self.data.shape[1:]), dtype=torch.float32, device=dev)
            op = None
            if self.emit_operand:
                from .engine import get_precision, tp_of
                N, T = self.data.shape[1], self.data.shape[2]
                op16 = torch.float16 if get_precision() == "fp16" else torch.bfloat16
                op = torch.empty(1, N, B, tp_of(T, "bf16"), dtype=op16, device=dev)
            K.assemble_batch(self.data, ids, table, inj, out, self.seed, self.draws, op)
            self.last_operand = op                        # Trainer.step(x, packed=loader.last_operand)
            self.draws += 1
            yield out


def create_augmented_dataloaders(x_data, batch_size, load_all=False, augmentation_config=None, val_split=0.2,
                                 num_workers=None, device=None):
    """Same signature and split semantics as the reference (augmentation.py:151-241): an 80/20 split by
    torch.randperm, shuffled augmented training batches, ordered plain validation batches.  The dataset is moved
    to the GPU once (the reference does the same for load_all=True, utils.py:41-43)."""
    dev = torch.device(device) if device is not None else torch.device("cuda")
    data = torch.as_tensor(x_data, dtype=torch.float32).to(dev).contiguous()
    n = data.shape[0]
    val_size = int(n * val_split)
    train_size = n - val_size
    perm = torch.randperm(n)
    train_idx, val_idx = perm[:train_size], perm[train_size:]
    train = B200AugmentedLoader(data, train_idx, batch_size, shuffle=True, augment=True)
    val = B200AugmentedLoader(data, val_idx, batch_size, shuffle=False, augment=False)
    return train, val
