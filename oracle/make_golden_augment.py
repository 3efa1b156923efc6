"""TEST INFRASTRUCTURE ONLY - golden batches of the reference's augmenting DataLoader.

Runs the UNMODIFIED /root/reference/modules/augmentation.py::create_augmented_dataloaders on a small synthetic
dataset with fixed seeds (python `random`, numpy, torch) and records every batch of two training epochs and one
validation epoch, together with the Gaussian noise tensors torch.randn_like handed out (so that the engine's fused
kernel can be fed the same noise).  Output: tests/golden/augment_toy.pt.  Usage: python oracle/make_golden_augment.py
"""
import importlib
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

P, N, T, BATCH = 11, 6, 16, 4
SEEDS = dict(python=5, numpy=6, torch=7)


def load_reference_augmentation():
    ref_import._install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    try:
        return importlib.import_module("modules.augmentation")
    finally:
        sys.path.remove(ref_import.REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def main():
    aug = load_reference_augmentation()
    g = torch.Generator().manual_seed(99)
    data = (torch.rand(P, N, T, generator=g) * 1.4 - 0.7).numpy().astype(np.float32)
    random.seed(SEEDS["python"])
    np.random.seed(SEEDS["numpy"])
    torch.manual_seed(SEEDS["torch"])
    train_dl, val_dl = aug.create_augmented_dataloaders(data, BATCH, load_all=True, augmentation_config=None)
    noise_gen = torch.Generator().manual_seed(1234)
    noise_log = []
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):
        e = torch.randn(t.shape, generator=noise_gen)
        noise_log.append(e)
        return e.to(t.device, t.dtype)

    record = dict(data=torch.from_numpy(data), batch=BATCH, seeds=SEEDS, train_epochs=[], val_epoch=[])
    torch.randn_like = fake_randn_like
    aug.torch.randn_like = fake_randn_like
    try:
        for _ in range(2):
            batches = []
            for x in train_dl:
                # noise tensors drawn while this batch was assembled, in sample order
                batches.append(dict(x=x.clone(), noise=[e.clone() for e in noise_log]))
                noise_log.clear()
            record["train_epochs"].append(batches)
        for x in val_dl:
            record["val_epoch"].append(x.clone())
    finally:
        torch.randn_like = orig
        aug.torch.randn_like = orig
    out = os.path.join(ROOT, "tests", "golden", "augment_toy.pt")
    torch.save(record, out)
    n_noise = sum(len(b["noise"]) for ep in record["train_epochs"] for b in ep)
    print("wrote", out, "train batches/epoch", len(record["train_epochs"][0]), "val batches", len(record["val_epoch"]),
          "noisy samples", n_noise)


if __name__ == "__main__":
    main()
