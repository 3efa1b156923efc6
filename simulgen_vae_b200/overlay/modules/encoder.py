"""Drop-in for /root/reference/modules/encoder.py (B200 engine overlay).

Parameter tree and forward signature as in the reference (encoder.py:14-167); the arithmetic is one
autograd Function over the engine's CUDA kernels (simulgen_vae_b200.engine.EncoderFn)."""
import torch.nn as nn

from modules.common import *  # noqa: F401,F403  (the reference re-exports common's names the same way)
from modules.common import _EngineBlock, conv_gn_gelu


class ConvBlock(_EngineBlock):
    """Conv k1 - GN - GELU (+ Conv k3 - GN - GELU when not small), reference encoder.py:14-58."""
    _sg_kind = "plain"

    def __init__(self, in_channel, out_channel, small):
        super().__init__()
        layers = conv_gn_gelu(in_channel, out_channel, 1)
        if not small:
            layers += conv_gn_gelu(out_channel, out_channel, 3)
        self._seq = nn.Sequential(*layers)


class EncoderBlock(_EngineBlock):
    """Chain of ConvBlocks over consecutive channel widths (reference encoder.py:60-94)."""
    _sg_kind = "chain"

    def __init__(self, channels, small):
        super().__init__()
        self.channels = channels
        self.module_list = nn.ModuleList(
            [ConvBlock(cin, cout, small) for cin, cout in zip(channels[:-1], channels[1:])])


class Encoder(nn.Module):
    """Hierarchical encoder (reference encoder.py:96-167): per level a ConvBlock and a residual block,
    a Linear(C*T -> hierarchical_dim) skip code per level and a final Linear(C*T -> 2*z_dim)."""

    def __init__(self, z_dim, hierarchical_dim, num_filter_enc, num_node, num_time, small):
        super().__init__()
        widths = [num_node] + list(num_filter_enc)
        self.encoder_blocks = nn.ModuleList(
            [EncoderBlock([widths[i], widths[i + 1]], small) for i in range(len(num_filter_enc))])
        self.encoder_residual_blocks = nn.ModuleList(
            [EncoderResidualBlock(c, c, small) for c in num_filter_enc])
        self.z_dim = z_dim
        self.num_filter_enc = num_filter_enc
        self.xs_linear = nn.ModuleList(
            [nn.Linear(c * num_time, int(hierarchical_dim)) for c in num_filter_enc])
        self.last_x_linear = nn.Linear(num_filter_enc[-1] * num_time, 2 * z_dim)
        self.small = small

    def _run(self, x, capture=None):
        """(last [B, 2*z_dim], xs list) - `last` keeps mu|log_var fused for the reparam kernel."""
        from simulgen_vae_b200 import engine
        engine.note_grad_mode()
        outs = engine.EncoderFn.apply(self, capture, x, *self.parameters())
        return outs[0], list(outs[1:])

    def forward(self, x):
        last, xs = self._run(x)
        return last[:, :self.z_dim], last[:, self.z_dim:], xs
