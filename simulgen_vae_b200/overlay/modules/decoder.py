"""Drop-in for /root/reference/modules/decoder.py (B200 engine overlay).

Parameter tree, `forward(z, xs, mode, freeze_level)` and `reparameterize(mu, std)` as in the reference
(decoder.py:17-223); the arithmetic is one autograd Function over the engine's CUDA kernels
(simulgen_vae_b200.engine.DecoderFn)."""
import torch
import torch.nn as nn
from torch.nn import functional as F  # noqa: F401

from modules.common import *  # noqa: F401,F403
from modules.common import _EngineBlock, gn_groups
from modules.losses import kl, kl_2  # noqa: F401


class UpsampleBlock(_EngineBlock):
    """ConvTranspose1d(k3, s1, p1) - GELU, no norm (reference decoder.py:17-45)."""
    _sg_kind = "upsample"

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self._seq = nn.Sequential(nn.ConvTranspose1d(in_channel, out_channel, kernel_size=3, padding=1), nn.GELU())


class DecoderBlock(_EngineBlock):
    """Chain of UpsampleBlocks (reference decoder.py:47-82)."""
    _sg_kind = "chain"

    def __init__(self, channels, small):
        super().__init__()
        self.channels = channels
        self.module_list = nn.ModuleList(
            [UpsampleBlock(cin, cout) for cin, cout in zip(channels[:-1], channels[1:])])


def _latent_to_sequence(dim, channels, num_time):
    """Linear(d, d*T) - Unflatten - Conv k5 - GN - GELU (reference decoder.py:131-148)."""
    return nn.Sequential(
        nn.Linear(dim, dim * num_time),
        nn.Unflatten(1, (dim, num_time)),
        nn.Conv1d(dim, channels, kernel_size=5, padding=2),
        nn.GroupNorm(gn_groups(channels), channels),
        nn.GELU(),
    )


def _condition(width_in, width_out, small):
    """ResidualBlock - GELU - Conv k3 (reference decoder.py:150-166)."""
    return nn.Sequential(ResidualBlock(width_in, small), nn.GELU(),
                         nn.Conv1d(width_in, width_out, kernel_size=3, padding=1))


class Decoder(nn.Module):
    """Hierarchical decoder (reference decoder.py:84-216)."""

    def __init__(self, z_dim, hierarchical_dim, num_filter_dec, num_node, num_time, batch_size, small):
        super().__init__()
        f = list(num_filter_dec)
        levels = range(len(f) - 1)
        self.decoder_blocks = nn.ModuleList([DecoderBlock([f[i], f[i + 1]], small) for i in levels])
        self.decoder_residual_blocks = nn.ModuleList([DecoderResidualBlock(f[i + 1], small) for i in levels])
        self.recon = nn.Sequential(
            nn.Conv1d(f[-1], num_node, kernel_size=1),
            nn.GroupNorm(gn_groups(num_node), num_node),
            nn.Tanh(),
        )
        self.zs = []
        self.num_filter_dec = num_filter_dec
        self.num_time = num_time
        self.sequence_start = nn.ModuleList([_latent_to_sequence(z_dim, f[0], num_time)])
        self.xs_sequence = nn.ModuleList([_latent_to_sequence(hierarchical_dim, f[i + 1], num_time) for i in levels])
        self.condition_z = nn.ModuleList([_condition(f[i + 1], 2 * f[i + 1], small) for i in levels])
        self.condition_xz = nn.ModuleList([_condition(2 * f[i + 1], 2 * f[i + 1], small) for i in levels])
        self.small = small

    def _run(self, z, xs, x=None, lossfun="MSE", mode="random", capture=None):
        from simulgen_vae_b200 import engine
        n_levels = len(self.decoder_residual_blocks) - 1
        if xs is None or len(xs) < n_levels:
            # the reference cannot run this either: z keeps its [B, z_dim] shape and torch.add fails
            # at the second level (decoder.py:179)
            raise RuntimeError("Decoder.forward needs xs with at least %d entries" % n_levels)
        xs = list(xs)
        engine.note_grad_mode()
        outs = engine.DecoderFn.apply(self, capture, lossfun, mode, len(xs), z, *xs, x, *self.parameters())
        x_hat = outs[0]
        if x is not None:
            return x_hat, outs[1], outs[2], list(outs[3:])
        return x_hat, None, None, list(outs[1:])

    def forward(self, z, xs=None, mode="random", freeze_level=-1):
        if freeze_level >= 0:
            from simulgen_vae_b200 import engine
            engine.set_freeze_level(freeze_level)            # decoder.py:202-207: cached / replayed latents in self.zs
        x_hat, _, _, kl_losses = self._run(z, xs, mode=mode)
        return x_hat, kl_losses


def reparameterize(mu, std):
    """z = mu + eps * clamp(std, 1e-8, 10) (reference decoder.py:218-223).  Standalone helper for
    external callers; VAE.forward uses the fused kernel instead."""
    from simulgen_vae_b200 import engine
    std = torch.clamp(std, min=1e-8, max=10.0)
    eps = engine.draw_eps(tuple(std.shape), std.device) if std.is_cuda else torch.randn_like(std)
    return mu + eps * std
