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

    def __init__(self, data, indices, batch_size, shuffle, augment, seed=0, rank=0, world=1, group=None):
        """batch_size is the PER-RANK batch (SimulGen-VAE.py:172 already divides Batch_size by the world size).  world > 1:
        every global batch of batch_size * world samples is split rank-major - rank r assembles positions
        [r * batch_size, (r + 1) * batch_size) of it - from ONE sampling order and ONE set of augmentation decisions
        (rank 0 draws them once per epoch, exactly as a single process would, and broadcasts them), so data-parallel
        training sees the batches of the single-process run at the global batch size.  The reference has no sampler at
        all: each rank would draw its own shuffle of the whole dataset (augmentation.py:182-187,226-232)."""
        self.data = data                                  # [P, N, T] fp32 CUDA, contiguous
        self.indices = torch.as_tensor(indices, dtype=torch.int64)
        self.batch_size, self.shuffle, self.augment = batch_size, shuffle, augment
        self.rank, self.world, self.group = int(rank), int(world), group
        self.dataset_len = data.shape[0]
        self._index_loader = DataLoader(Subset(_IndexDataset(self.dataset_len), self.indices.tolist()),
                                        batch_size=batch_size * self.world, shuffle=shuffle, num_workers=0)
        self.seed, self.draws = seed, 0
        self.injected_noise = None                        # tests: callable(batch_position, shape) -> noise tensor
        self.last_decisions = None
        self.emit_operand = False                         # also produce the packed 16-bit conv operand
        self.yield_packed = False                         # yield engine.PackedBatch (operand only, no fp32 batch)
        self.prefetch = False                             # assemble batch i + 1 on a side stream while batch i is consumed
        self.last_operand = None
        self._stream = None
        self._ring = {}

    def __len__(self):
        n = len(self._index_loader)
        if self.world > 1 and len(self.indices) % (self.batch_size * self.world) % self.world:
            n -= 1                                        # a ragged last global batch that cannot be split evenly is dropped
        return n

    def _decide(self, idx):
        B = len(idx)
        if self.augment:
            return draw_decisions(idx, self.dataset_len)
        return (np.zeros(B, np.float32), np.ones(B, np.float32), np.full(B, -1, np.int64), np.ones(B, np.float32),
                np.zeros(B, np.float32))

    def _global_batches(self):
        """(indices, decisions) of this rank's share of every global batch of the epoch."""
        if self.world == 1:
            for idx in self._index_loader:
                idx = idx.reshape(-1)
                yield idx, self._decide(idx.tolist())
            return
        dist = torch.distributed
        n, W = len(self.indices), self.world
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        on_dev = dist.get_backend(self.group) == "nccl"
        plan = torch.zeros(6, n, dtype=torch.float64)
        if self.rank == 0:
            # the whole epoch at once, in the order a single process would draw it batch by batch
            order = torch.cat([idx.reshape(-1) for idx in self._index_loader])
            cols = [self._decide(order[i:i + self.batch_size * W].tolist()) for i in range(0, n, self.batch_size * W)]
            dec = [np.concatenate([c[j] for c in cols]) for j in range(5)]
            plan = torch.from_numpy(np.stack([order.numpy().astype(np.float64)] + [d.astype(np.float64) for d in dec]))
        if on_dev:
            plan = plan.to(self.data.device)
        dist.broadcast(plan, src=src, group=self.group)
        plan = plan.cpu()
        Bg = self.batch_size * W
        for g0 in range(0, n, Bg):
            m = min(Bg, n - g0)
            if m % W:
                break                                     # ragged tail that cannot be split evenly
            per = m // W
            sl = slice(g0 + self.rank * per, g0 + (self.rank + 1) * per)
            idx = plan[0, sl].to(torch.int64)
            yield idx, (plan[1, sl].numpy().astype(np.float32), plan[2, sl].numpy().astype(np.float32),
                        plan[3, sl].numpy().astype(np.int64), plan[4, sl].numpy().astype(np.float32),
                        plan[5, sl].numpy().astype(np.float32))

    def _assemble(self, idx, decisions, slot=None):
        """Launch sg_assemble_batch for one batch on the current stream; returns (what the loader yields, operand).
        slot: ring slot whose preallocated output buffers are used (prefetch mode), else fresh allocations."""
        noise, scale, other, lam, om = decisions
        dev = self.data.device
        B = idx.numel()
        self.last_decisions = dict(index=idx.clone(), noise=noise, scale=scale, other=other, lam=lam)
        table = torch.from_numpy(np.stack([noise, scale, lam, om]).astype(np.float32)).to(dev, non_blocking=True)
        ids = torch.stack([idx, torch.from_numpy(other)]).to(torch.int32).to(dev, non_blocking=True)
        inj = None
        if self.injected_noise is not None and (noise > 0).any():
            inj = self.injected_noise(noise, tuple(self.data.shape[1:])).to(dev)
        from .engine import PackedBatch, get_precision, loss_target, tp_of
        N, T = self.data.shape[1], self.data.shape[2]
        packed_only = self.yield_packed and loss_target(T) == "operand"
        op16 = torch.float16 if get_precision() == "fp16" else torch.bfloat16
        want_op = self.emit_operand or packed_only
        if slot is not None:
            ring = self._ring.setdefault(slot, {})
            key = (B, packed_only, want_op, op16)
            if ring.get("key") != key:                    # (re)allocate this slot's buffers: first use or a ragged batch
                ring.clear()
                ring["key"] = key
                ring["out"] = None if packed_only else torch.empty((B, N, T), dtype=torch.float32, device=dev)
                ring["op"] = torch.empty(1, N, B, tp_of(T, "bf16"), dtype=op16, device=dev) if want_op else None
            out, op = ring["out"], ring["op"]
        else:
            out = None if packed_only else torch.empty((B, N, T), dtype=torch.float32, device=dev)
            op = torch.empty(1, N, B, tp_of(T, "bf16"), dtype=op16, device=dev) if want_op else None
        K.assemble_batch(self.data, ids, table, inj, out, self.seed, self.draws, op,
                         blocks_per_sm=2 if slot is not None else 0)
        self.draws += 1
        return (PackedBatch(op, T) if packed_only else out), op, (ids, table, inj)

    def __iter__(self):
        dev = self.data.device
        if not (self.prefetch and dev.type == "cuda"):
            for idx, decisions in self._global_batches():
                batch, op, _ = self._assemble(idx, decisions)
                self.last_operand = op                    # Trainer.step(x, packed=loader.last_operand)
                yield batch
            return
        # One batch ahead: the gather / augmentation kernel of batch i + 1 is HBM-bound and runs on its own stream
        # underneath the (tensor-core-bound) training step of batch i, with a small resident footprint (2 thread blocks
        # per SM) so that the step's persistent GEMM CTAs still fit next to it.  Output buffers come from a ring of 3
        # slots owned by the loader (no allocator traffic, no cudaMalloc in steady state): a yielded batch stays valid
        # until the second-next batch is requested.  Slot reuse is ordered by events: the kernel that refills the slot of
        # batch i - 2 waits for the marker the consumer's stream recorded when batch i - 1 was handed over, i.e. for
        # everything that consumed batch i - 2.
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        side = self._stream
        markers = {}
        state = {"n": 0}

        def launch(idx, decisions):
            i = state["n"]
            state["n"] += 1
            with torch.cuda.stream(side):
                if i - 2 in markers:
                    side.wait_event(markers.pop(i - 2))
                batch, op, keep = self._assemble(idx, decisions, slot=i % 3)
                ev = torch.cuda.Event()
                ev.record(side)
            return i, batch, op, ev, keep

        def hand_over(item):
            i, batch, op, ev, keep = item
            cur = torch.cuda.current_stream(dev)
            m = torch.cuda.Event()
            m.record(cur)                                 # everything issued so far (the consumers of batch i - 1) ...
            markers[i] = m                                # ... must finish before slot (i - 1) % 3 is written for batch i + 2
            cur.wait_event(ev)
            self.last_operand = op
            return batch

        pending = None
        for idx, decisions in self._global_batches():
            nxt = launch(idx, decisions)
            if pending is not None:
                yield hand_over(pending)
            pending = nxt
        if pending is not None:
            yield hand_over(pending)


def create_augmented_dataloaders(x_data, batch_size, load_all=False, augmentation_config=None, val_split=0.2,
                                 num_workers=None, device=None, process_group=None):
    """Same signature and split semantics as the reference (augmentation.py:151-241): an 80/20 split by
    torch.randperm, shuffled augmented training batches, ordered plain validation batches.  The dataset is moved
    to the GPU once (the reference does the same for load_all=True, utils.py:41-43).

    Under torch.distributed (SimulGen-VAE.py --use_ddp; `batch_size` is then already the per-rank share,
    SimulGen-VAE.py:168-174) the split is rank 0's, broadcast to every rank, and both loaders hand rank r its
    disjoint 1/W slice of every global batch (see B200AugmentedLoader)."""
    dist = torch.distributed
    ddp = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(process_group) if ddp else 0
    world = dist.get_world_size(process_group) if ddp else 1
    if device is None:
        import os
        device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if ddp else "cuda"
    dev = torch.device(device)
    data = torch.as_tensor(x_data, dtype=torch.float32).to(dev).contiguous()
    n = data.shape[0]
    val_size = int(n * val_split)
    train_size = n - val_size
    perm = torch.randperm(n)
    if world > 1:
        src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
        on_dev = dist.get_backend(process_group) == "nccl"
        perm = perm.to(dev) if on_dev else perm
        dist.broadcast(perm, src=src, group=process_group)
        perm = perm.cpu()
    train_idx, val_idx = perm[:train_size], perm[train_size:]
    train = B200AugmentedLoader(data, train_idx, batch_size, shuffle=True, augment=True, rank=rank, world=world,
                                group=process_group)
    val = B200AugmentedLoader(data, val_idx, batch_size, shuffle=False, augment=False, rank=rank, world=world,
                              group=process_group)
    return train, val
