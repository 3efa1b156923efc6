"""TEST INFRASTRUCTURE ONLY - plain PyTorch fp32 restatement of the reference's hot path.

This file is the *oracle* for the parity tests (`tests/`), `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  Nothing under `simulgen_vae_b200/` may
import it: the product path is hand-written CUDA only and fails loudly without its extension.

It restates, as pure functions over a reference-layout `state_dict`, exactly the arithmetic of

  * /root/reference/modules/VAE_network.py:79-117   (VAE.forward)
  * /root/reference/modules/encoder.py:29-46,146-167 (ConvBlock, Encoder.forward)
  * /root/reference/modules/decoder.py:27-33,106-223 (UpsampleBlock, Decoder.forward, reparameterize)
  * /root/reference/modules/common.py:78-162         (Residual / EncoderResidual / DecoderResidual blocks)
  * /root/reference/modules/losses.py:8-48           (kl, kl_2)
  * /root/reference/modules/train.py:18-41,144-150   (WarmupKLLoss, loss assembly)
  * torch/nn/utils/spectral_norm.py:62-114           (hook-style spectral norm, one power iteration)

Pinning: `oracle/make_golden.py` runs the real reference (imported from /root/reference with
`oracle/ref_import.py`) and this restatement on the same seeds and asserts they agree to fp32
round-off before writing `tests/golden/*.pt`; `tests/test_oracle.py` re-checks the restatement
against those committed fixtures (and against the live reference when the checkout is present).
The reference itself ships no tests or golden vectors, so the reference's own
outputs generated here are the only anchor.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SN_EPS = 1e-12          # torch.nn.utils.spectral_norm default eps
GN_EPS = 1e-5           # nn.GroupNorm default eps


def gn_groups(c: int) -> int:
    """encoder.py:35 / common.py:85 / decoder.py:119: min(8, max(1, C // 4))."""
    return min(8, max(1, c // 4))


# ----------------------------------------------------------------------------------------------
# spectral norm (torch/nn/utils/spectral_norm.py:62-114), functional
# ----------------------------------------------------------------------------------------------
def sn_weight(sd: Dict[str, torch.Tensor], prefix: str, training: bool, transposed: bool = False):
    """Return W_orig / sigma.  In training mode advance (u, v) by one power iteration first and
    store the new vectors back into `sd` (they are buffers: no gradient flows through them).
    `transposed` = ConvTranspose1d (spectral_norm uses dim=1 there, spectral_norm.py:329-333)."""
    w = sd[prefix + ".weight_orig"]
    u = sd[prefix + ".weight_u"]
    v = sd[prefix + ".weight_v"]
    wm = w
    if transposed:
        wm = wm.permute(1, 0, *range(2, wm.dim()))
    wm = wm.reshape(wm.shape[0], -1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS)
            sd[prefix + ".weight_u"] = u.clone()
            sd[prefix + ".weight_v"] = v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def conv1d_sn(sd, prefix, x, training):
    w = sn_weight(sd, prefix, training)
    k = w.shape[2]
    return F.conv1d(x, w, sd[prefix + ".bias"], padding=(k - 1) // 2)


def convT1d_sn(sd, prefix, x, training):
    w = sn_weight(sd, prefix, training, transposed=True)
    k = w.shape[2]
    return F.conv_transpose1d(x, w, sd[prefix + ".bias"], padding=(k - 1) // 2)


def linear_sn(sd, prefix, x, training):
    return F.linear(x, sn_weight(sd, prefix, training), sd[prefix + ".bias"])


def gn(sd, prefix, x):
    c = x.shape[1]
    return F.group_norm(x, gn_groups(c), sd[prefix + ".weight"], sd[prefix + ".bias"], GN_EPS)


def cgg_seq(sd, prefix, x, training, acts=None):
    """nn.Sequential of (Conv1d, GroupNorm, GELU) triples at indices 0-2, 3-5, ...
    (encoder.py:31-46, common.py:82-99,108-123,133-160)."""
    i = 0
    while f"{prefix}.{i}.weight_orig" in sd:
        x = conv1d_sn(sd, f"{prefix}.{i}", x, training)
        x = F.gelu(gn(sd, f"{prefix}.{i + 1}", x))
        if acts is not None:
            acts[f"{prefix}.{i + 2}"] = x
        i += 3
    return x


# ----------------------------------------------------------------------------------------------
# losses (losses.py:8-48)
# ----------------------------------------------------------------------------------------------
def kl(mu, log_var):
    log_var = torch.clamp(log_var, min=-30, max=30)
    var = torch.exp(log_var)
    loss = 0.5 * torch.sum(mu ** 2 + var - log_var - 1, dim=[1])
    return torch.mean(loss, dim=0)


def kl_2(delta_mu, delta_log_var, mu, log_var):
    log_var = torch.clamp(log_var, min=-30, max=30)
    delta_log_var = torch.clamp(delta_log_var, min=-30, max=30)
    var = torch.exp(log_var) + 1e-8
    delta_var = torch.exp(delta_log_var)
    loss = 0.5 * torch.sum(delta_var / var + (mu - delta_mu) ** 2 / var - delta_log_var + log_var - 1,
                           dim=[1, 2])
    return torch.mean(loss, dim=0)


def reparameterize(mu, std, eps):
    """decoder.py:218-223 with the eps draw made explicit."""
    std = torch.clamp(std, min=1e-8, max=10.0)
    return mu + eps * std


def recon_loss(kind: str, x_hat, x):
    """VAE_network.py:71-77: nn.MSELoss / L1Loss / SmoothL1Loss(beta=1) / HuberLoss(delta=1), mean
    reduction; unknown names fall back to MSE (`.get(self.lossfun, self.mse_loss)`)."""
    if kind == "MAE":
        return F.l1_loss(x_hat, x)
    if kind == "smoothL1":
        return F.smooth_l1_loss(x_hat, x)
    if kind == "Huber":
        return F.huber_loss(x_hat, x)
    return F.mse_loss(x_hat, x)


# ----------------------------------------------------------------------------------------------
# encoder / decoder / VAE
# ----------------------------------------------------------------------------------------------
def num_levels(sd) -> int:
    n = 0
    while f"encoder.encoder_blocks.{n}.module_list.0._seq.0.weight_orig" in sd:
        n += 1
    return n


def encoder_forward(sd, x, latent_dim: int, training: bool, acts=None):
    """encoder.py:146-167."""
    B = x.shape[0]
    xs = []
    h = x
    for i in range(num_levels(sd)):
        h = cgg_seq(sd, f"encoder.encoder_blocks.{i}.module_list.0._seq", h, training, acts)
        h = h + 0.1 * cgg_seq(sd, f"encoder.encoder_residual_blocks.{i}.seq", h, training, acts)
        if acts is not None:
            acts[f"encoder.level{i}"] = h
        xs.append(linear_sn(sd, f"encoder.xs_linear.{i}", h.reshape(B, -1), training))
    last = linear_sn(sd, "encoder.last_x_linear", h.reshape(B, -1), training)
    mu = last[:, :latent_dim]
    log_var = last[:, latent_dim:]
    return mu, log_var, xs[:-1][::-1]


def _latent_seq(sd, prefix, z, T, training):
    """decoder.py:131-148: Linear(d, d*T) -> Unflatten(1,(d,T)) -> Conv k5 -> GN -> GELU."""
    d = z.shape[1]
    h = linear_sn(sd, prefix + ".0", z, training).reshape(z.shape[0], d, T)
    h = conv1d_sn(sd, prefix + ".2", h, training)
    return F.gelu(gn(sd, prefix + ".3", h))


def _condition(sd, prefix, h, training):
    """decoder.py:150-166: ResidualBlock -> GELU -> Conv k3 (no norm)."""
    r = h + 0.1 * cgg_seq(sd, prefix + ".0._seq", h, training)
    return conv1d_sn(sd, prefix + ".2", F.gelu(r), training)


def decoder_forward(sd, z, xs, eps_list: Optional[List[torch.Tensor]], num_time: int, training: bool,
                    mode: str = "random", acts=None):
    """decoder.py:170-216 (freeze_level < 0, the only value any caller passes)."""
    n_blocks = 0
    while f"decoder.decoder_residual_blocks.{n_blocks}.seq.0.weight_orig" in sd:
        n_blocks += 1
    kl_losses = []
    eps_i = 0
    out = None
    for i in range(n_blocks):
        if i == 0:
            z_sample = _latent_seq(sd, "decoder.sequence_start.0", z, num_time, training)
        else:
            z_sample = out + z
        out = F.gelu(convT1d_sn(sd, f"decoder.decoder_blocks.{i}.module_list.0._seq.0", z_sample, training))
        if acts is not None:
            acts[f"decoder.up{i}"] = out
        out = out + 0.1 * cgg_seq(sd, f"decoder.decoder_residual_blocks.{i}.seq", out, training, acts)
        if acts is not None:
            acts[f"decoder.level{i}"] = out
        if i == n_blocks - 1:
            break
        mu, log_var = _condition(sd, f"decoder.condition_z.{i}", out, training).chunk(2, dim=1)
        if xs is not None:
            xs_sample = _latent_seq(sd, f"decoder.xs_sequence.{i}", xs[i], num_time, training)
            d_mu, d_lv = _condition(sd, f"decoder.condition_xz.{i}", torch.cat([xs_sample, out], dim=1),
                                    training).chunk(2, dim=1)
            kl_losses.append(kl_2(d_mu, d_lv, mu, log_var))
            mu = mu + d_mu
            log_var = torch.clamp(log_var + d_lv, min=-30, max=30)
            std = torch.exp(0.5 * log_var)
            if mode == "fix":
                std = std * 1e-10
            z = reparameterize(mu, std, eps_list[eps_i])
            eps_i += 1
            if acts is not None:
                acts[f"decoder.z{i}"] = z
    y = conv1d_sn(sd, "decoder.recon.0", out, training)
    x_hat = torch.tanh(gn(sd, "decoder.recon.1", y))
    return x_hat, kl_losses


def vae_forward(sd, x, eps_list: List[torch.Tensor], latent_dim: int, lossfun: str = "MSE",
                training: bool = True, acts=None):
    """VAE_network.py:79-117.  `eps_list` = [eps0 [B,latent], eps1 [B,C1,T], ...] in draw order.
    Returns (x_hat, recon_loss, [kl_main, *kl2], recon_loss_MSE)."""
    T = x.shape[2]
    mu, log_var, xs = encoder_forward(sd, x, latent_dim, training, acts)
    log_var = torch.clamp(log_var, min=-30, max=30)
    std = torch.exp(0.5 * log_var)
    z = reparameterize(mu, std, eps_list[0])
    if acts is not None:
        acts["mu"], acts["log_var"], acts["z"] = mu, log_var, z
        for i, t in enumerate(xs):
            acts[f"xs{i}"] = t
    x_hat, kl_losses = decoder_forward(sd, z, xs, eps_list[1:], T, training, acts=acts)
    rl = recon_loss(lossfun, x_hat, x)
    mse = F.mse_loss(x_hat, x)
    return x_hat, rl, [kl(mu, log_var)] + kl_losses, mse


# ----------------------------------------------------------------------------------------------
# train.py step semantics
# ----------------------------------------------------------------------------------------------
def warmup_beta(epoch: int, epochs: int, init_beta: float = 1e-4, beta_target: float = 1.0) -> float:
    """train.py:18-41,75-81."""
    start, end = int(epochs * 0.3), int(epochs * 0.8)
    if epoch < start:
        return init_beta
    if start <= epoch < end:
        return (epoch - start) * (beta_target - init_beta) / (end - start) + init_beta
    return beta_target


def total_loss(recon, kl_losses, alpha: float, beta: float):
    """train.py:144-150."""
    s = 0
    for l in kl_losses:
        s = s + l
    return recon * alpha + s * beta


def eps_shapes(cfg, batch: int):
    """Shapes of the randn_like draws per forward, in order (decoder.py:221)."""
    dec = list(cfg["enc"])[::-1]
    shapes = [(batch, cfg["latent_dim"])]
    for i in range(len(dec) - 2):
        shapes.append((batch, dec[i + 1], cfg["num_time"]))
    return shapes


def params_from_state_dict(sd, requires_grad=True):
    """Clone a state dict into leaf tensors (parameters get requires_grad, u/v buffers do not)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and not (k.endswith("weight_u") or k.endswith("weight_v")):
            t.requires_grad_(True)
        out[k] = t
    return out


def synthetic_field(P: int, N: int, T: int, seed: int = 1234, device="cpu", dtype=torch.float32):
    """SURVEY.md 8d generator: x[p,n,t] = 0.7 a_n sin(2 pi (f_p t/T + phi_n)) + 0.02 xi, clipped to
    +-0.7 (the range data_preprocess.py:90 produces), laid out [P, N, T] (SimulGen-VAE.py:282)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.rand(N, generator=g) * 0.8 + 0.2
    phi = torch.rand(N, generator=g)
    f = torch.rand(P, generator=g) * 3.5 + 0.5
    t = torch.arange(T, dtype=torch.float32) / T
    x = 0.7 * a[None, :, None] * torch.sin(2 * math.pi * (f[:, None, None] * t[None, None, :] + phi[None, :, None]))
    x = x + 0.02 * torch.randn(P, N, T, generator=g)
    return x.clamp_(-0.7, 0.7).to(device=device, dtype=dtype)
