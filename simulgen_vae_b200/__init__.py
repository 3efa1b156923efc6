"""simulgen_vae_b200 - B200-native engine for the SimulGen-VAE hot path.

`install_overlay()` puts `simulgen_vae_b200/overlay` at the front of sys.path so that the reference's
own callers (`from modules.VAE_network import VAE` in train.py:10 / SimulGen-VAE.py:74, `from
modules.common import initialize_weights_He, add_sn`, ...) bind to the engine's drop-in modules; the
reference's `modules/` is a namespace package (no __init__.py), so the remaining files
(train.py, utils.py, ...) still load from the reference checkout.
"""
import os
import sys

from .engine import DEFAULT_PRECISION, PackedBatch, fixed_eps, loss_target, set_loss_target, get_precision, set_precision, set_sample_offset, tp_of  # noqa: F401

OVERLAY_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay")
OVERLAY_TRAIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay_train")


def install_overlay(train=None):
    """Make `import modules.{VAE_network,encoder,decoder,common,losses}` resolve to the engine.  train=True also
    shadows `modules.train` with the Trainer-based driver (simulgen_vae_b200.train_loop); by default the reference's own
    train.py keeps running on top of the overlaid model modules (train=None keeps the current choice)."""
    if train is None:
        train = OVERLAY_TRAIN_DIR in sys.path
    for d in (OVERLAY_DIR, OVERLAY_TRAIN_DIR):
        if d in sys.path:
            sys.path.remove(d)
    sys.path.insert(0, OVERLAY_DIR)
    if train:
        sys.path.insert(0, OVERLAY_TRAIN_DIR)
        sys.modules.pop("modules.train", None)
    stale = [k for k, m in sys.modules.items()
             if (k == "modules" or k.startswith("modules.")) and OVERLAY_DIR not in (getattr(m, "__file__", None) or OVERLAY_DIR)]
    for k in stale:
        if k.split(".")[-1] in ("modules", "VAE_network", "encoder", "decoder", "common", "losses"):
            del sys.modules[k]
    return OVERLAY_DIR


def load_vae_class():
    install_overlay()
    from modules.VAE_network import VAE
    return VAE
