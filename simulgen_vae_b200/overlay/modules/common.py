"""Drop-in for /root/reference/modules/common.py (B200 engine overlay).

Same public names and state-dict layout; the arithmetic runs in simulgen_vae_b200 (CUDA kernels).
The block classes here are parameter containers: their layers are real nn.Conv1d / nn.GroupNorm
objects registered under the reference's attribute names so that `model.apply(initialize_weights_He)`,
`model.apply(add_sn)`, `state_dict()` (bias, weight_orig, weight_u, weight_v per wrapped layer) and
whole-module pickles behave exactly as with the reference (common.py:15-59, train.py:71-72,252-253).
"""
import math  # noqa: F401  (the reference module exports it through `from modules.common import *`)

import numpy as np  # noqa: F401
import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

_SN_TYPES = (nn.Conv1d, nn.ConvTranspose1d, nn.Conv2d, nn.ConvTranspose2d, nn.Linear)
_CONV_TYPES = (nn.Conv1d, nn.ConvTranspose1d, nn.Conv2d, nn.ConvTranspose2d)


def add_sn(m):
    """Register hook-style spectral normalisation on conv / linear layers (reference common.py:15-37).

    Registration (weight_orig / weight_u / weight_v, state-dict hooks, the u, v draws) is delegated to
    torch.nn.utils.spectral_norm so checkpoints and RNG order are identical to the reference.  Inside
    the engine the layer is never *called*: the power iteration, W/sigma and their backward run in
    sg_sn_power_iter / sg_sn_pack_weight / sg_sn_weight_grad on the registered tensors."""
    if not isinstance(m, _SN_TYPES):
        return m
    if m.weight.numel() == 0:
        print(f'Warning: Cannot apply spectral normalization to {type(m).__name__} - weight tensor is empty')
        return m
    return spectral_norm(m)


def initialize_weights_He(m):
    """Kaiming-uniform weights, zero biases (reference common.py:39-59)."""
    if isinstance(m, _CONV_TYPES):
        nn.init.kaiming_uniform_(m.weight.data, nonlinearity='relu')
        if m.bias is not None:
            nn.init.constant_(m.bias.data, 0)
    elif isinstance(m, nn.Linear):
        nn.init.kaiming_uniform_(m.weight.data)
        nn.init.constant_(m.bias.data, 0)


def gn_groups(channels):
    return min(8, max(1, channels // 4))


def conv_gn_gelu(cin, cout, k):
    """The (Conv1d, GroupNorm, GELU) triple every block of the model is made of."""
    return [nn.Conv1d(cin, cout, kernel_size=k, padding=(k - 1) // 2), nn.GroupNorm(gn_groups(cout), cout), nn.GELU()]


class Swish(nn.Module):
    """x * sigmoid(x) - exported by the reference (common.py:63-76), unused by the model."""

    def forward(self, x):
        return x * torch.sigmoid(x)


class _EngineBlock(nn.Module):
    """Base of the container blocks: standalone calls run the block through the engine."""
    _sg_kind = None

    def forward(self, x):
        from simulgen_vae_b200 import blocks
        return blocks.run_block(self, x)


class ResidualBlock(_EngineBlock):
    """x + 0.1 * seq(x), seq = 1 (small) or 2 conv-GN-GELU triples of width dim (common.py:78-102)."""
    _sg_kind = "residual"

    def __init__(self, dim, small):
        super().__init__()
        layers = conv_gn_gelu(dim, dim, 3)
        if not small:
            layers += conv_gn_gelu(dim, dim, 3)
        self._seq = nn.Sequential(*layers)


class EncoderResidualBlock(_EngineBlock):
    """Same arithmetic as ResidualBlock, registered under `seq` (common.py:104-125)."""
    _sg_kind = "residual"

    def __init__(self, input, dim, small):
        super().__init__()
        layers = conv_gn_gelu(input, input, 3)
        if not small:
            layers += conv_gn_gelu(input, input, 3)
        self.seq = nn.Sequential(*layers)


class DecoderResidualBlock(_EngineBlock):
    """x + 0.1 * seq(x) with a x5 channel expansion (common.py:127-162)."""
    _sg_kind = "residual"

    def __init__(self, input, small):
        super().__init__()
        wide = input * 5
        if small:
            plan = [(input, wide, 1), (wide, wide, 5), (wide, input, 1)]
        else:
            plan = [(input, input, 1), (input, wide, 5), (wide, wide, 5), (wide, input, 1)]
        layers = []
        for cin, cout, k in plan:
            layers += conv_gn_gelu(cin, cout, k)
        self.seq = nn.Sequential(*layers)
