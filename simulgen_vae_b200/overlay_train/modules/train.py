"""Drop-in for /root/reference/modules/train.py (B200 engine overlay).

`train(epochs, batch_size, train_dataloader, val_dataloader, LR, num_filter_enc, num_filter_dec, num_node, latent_dim,
hierarchical_dim, num_time, alpha, lossfun, small, load_all, debug_mode=0)` (train.py:49) keeps the reference's
signature, schedules, checkpoints and return value; the batch loop runs on simulgen_vae_b200.trainer.Trainer (fused
optimiser, no per-step host synchronisation, NCCL data parallelism when torch.distributed is initialised).
This overlay directory is opt-in: `install_overlay(train=True)`; by default the reference's own train.py keeps running
on top of the overlaid model modules."""
from simulgen_vae_b200.train_loop import WarmupKLLoss, print_gpu_mem_checkpoint, train  # noqa: F401
