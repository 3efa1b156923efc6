"""Drop-in for /root/reference/modules/augmentation.py (B200 engine overlay).

`create_augmented_dataloaders(x_data, batch_size, load_all, augmentation_config, val_split, num_workers)` keeps
the reference's signature, split and sampling semantics (augmentation.py:151-241) and its augmentation recipe
(augmentation.py:26-38,57-124) but assembles every batch with one CUDA kernel from a GPU-resident dataset
(simulgen_vae_b200.augment -> sg_assemble_batch).  With the same python / numpy / torch seeds the batches are
bit-identical to the reference's apart from the values of the Gaussian noise (Philox instead of torch.randn_like)."""
from simulgen_vae_b200.augment import B200AugmentedLoader, DEFAULTS, create_augmented_dataloaders, draw_decisions  # noqa: F401
