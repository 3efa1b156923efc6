"""Drop-in for /root/reference/modules/VAE_network.py (B200 engine overlay).

Same constructor, attributes and `forward(x) -> (x_hat, recon_loss, [kl_main, *kl_2], recon_loss_MSE)`
as the reference (VAE_network.py:33-164).  The forward is three autograd Functions over hand-written
sm_100a kernels: encoder, main-latent reparameterisation + KL, decoder + reconstruction losses."""
import torch
import torch.nn as nn

from modules.encoder import Encoder
from modules.decoder import Decoder, reparameterize  # noqa: F401
from modules.losses import kl  # noqa: F401
from modules.common import add_sn  # noqa: F401


class VAE(nn.Module):
    def __init__(self, latent_dim, hierarchical_dim, num_filter_enc, num_filter_dec, num_node, num_time,
                 lossfun='MSE', batch_size=1, small=False, use_checkpointing=False):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder = Encoder(latent_dim, hierarchical_dim, num_filter_enc, num_node, num_time, small)
        self.decoder = Decoder(latent_dim, hierarchical_dim, num_filter_dec, num_node, num_time, batch_size, small)
        self.lossfun = lossfun
        self.use_checkpointing = False          # the reference forces it off (VAE_network.py:68)
        # kept for attribute / pickle parity (VAE_network.py:71-77); the engine fuses the reductions
        self.loss_functions = {'MSE': nn.MSELoss(), 'MAE': nn.L1Loss(), 'smoothL1': nn.SmoothL1Loss(),
                               'Huber': nn.HuberLoss()}
        self.mse_loss = nn.MSELoss()

    def forward(self, x, _capture=None):
        try:
            from simulgen_vae_b200 import engine
            last, xs = self.encoder._run(x, _capture)
            # the packed 16-bit operand of x the encoder just consumed doubles as the loss target (engine.loss_target)
            engine.set_loss_operand(engine.take_last_packed())
            eps0 = engine.draw_eps((x.shape[0], self.latent_dim), x.device)
            z, kl_main = engine.ReparamMainFn.apply(last, eps0)
            lossfun = self.lossfun if self.lossfun in self.loss_functions else 'MSE'
            x_hat, recon_loss, recon_loss_MSE, kl_losses = self.decoder._run(z, xs, x=x, lossfun=lossfun,
                                                                            capture=_capture)
            return x_hat, recon_loss, [kl_main] + kl_losses, recon_loss_MSE
        except RuntimeError as e:
            print(f"Error in VAE forward pass: {e}")
            raise

    def compile_model(self, mode='max-autotune'):
        """The reference wraps encoder/decoder in torch.compile (VAE_network.py:123-152) and train.py
        calls it with 'none'.  The engine is already one fused CUDA path per sub-network, so every
        mode is a no-op here."""
        return

    def to(self, *args, **kwargs):
        device = args[0] if args else kwargs.get('device', None)
        if device:
            print("Moving model to device (channels_last disabled for compilation compatibility)")
        return super().to(*args, **kwargs)
