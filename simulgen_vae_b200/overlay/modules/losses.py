"""Drop-in for /root/reference/modules/losses.py.

Inside VAE.forward the KL terms are produced by the fused CUDA kernels (sg_reparam_main_fwd,
sg_kl2_reparam_fwd).  These standalone functions keep the reference's public API (losses.py:8-53)
for callers that evaluate a KL on tensors of their own; they are not on the training hot path."""
import torch


def kl(mu, log_var):
    """mean_b( 0.5 * sum_d( mu^2 + exp(lv) - lv - 1 ) ), lv clamped to [-30, 30] (losses.py:8-32)."""
    lv = torch.clamp(log_var, min=-30, max=30)
    return torch.mean(0.5 * torch.sum(mu * mu + torch.exp(lv) - lv - 1, dim=[1]), dim=0)


def kl_2(delta_mu, delta_log_var, mu, log_var):
    """KL between the conditional posterior and prior of a hierarchical level (losses.py:34-48);
    note the reference's (mu - delta_mu)^2 term, kept as written."""
    lv = torch.clamp(log_var, min=-30, max=30)
    dlv = torch.clamp(delta_log_var, min=-30, max=30)
    var = torch.exp(lv) + 1e-8
    integrand = torch.exp(dlv) / var + (mu - delta_mu) ** 2 / var - dlv + lv - 1
    return torch.mean(0.5 * torch.sum(integrand, dim=[1, 2]), dim=0)


def log_sum_exp(x):
    m2 = torch.max(x, dim=1, keepdim=True)[0]
    return m2.unsqueeze(1) + torch.log(torch.sum(torch.exp(x - m2), dim=1))
