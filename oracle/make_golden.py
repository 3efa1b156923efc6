"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

For every case it (1) builds the reference VAE exactly like modules/train.py:65-72, (2) runs
forward + backward with a fixed eps stream (torch.randn_like patched), (3) runs the oracle
restatement (oracle/vae_oracle.py) on the same state dict / inputs and asserts agreement to fp32
round-off - this is what pins the oracle - and (4) stores inputs and the reference's outputs.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, vae_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: cfg
    "toy3_small_mse": dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20,
                           small=True, batch=4, lossfun="MSE"),
    "toy4_small_huber": dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 16, 8], num_node=48, num_time=12,
                             small=True, batch=3, lossfun="Huber"),
    "toy4_large_mae": dict(latent_dim=16, hierarchical_dim=8, enc=[16, 16, 8, 8], num_node=40, num_time=10,
                           small=False, batch=2, lossfun="MAE"),
    "toy3_small_smoothl1": dict(latent_dim=32, hierarchical_dim=8, enc=[16, 8, 8], num_node=32, num_time=8,
                                small=True, batch=2, lossfun="smoothL1"),
}
ALPHA = 1.0e6
BETA = 1.0e-4


def rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run_case(name, cfg):
    torch.manual_seed(0)
    model = ref_import.build_reference_vae(cfg, seed=7)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B, N, T = cfg["batch"], cfg["num_node"], cfg["num_time"]
    x = O.synthetic_field(B, N, T, seed=11)
    g = torch.Generator().manual_seed(5)
    eps = [torch.randn(s, generator=g) for s in O.eps_shapes(cfg, B)]

    # ---- reference, training mode -------------------------------------------------------------
    model.train(True)
    with ref_import.patched_randn_like(eps):
        x_hat, rl, kls, mse = model(x)
    loss = rl * ALPHA + sum(kls) * BETA
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    sd1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref_out = dict(x_hat=x_hat.detach().clone(), recon=rl.detach().clone(), mse=mse.detach().clone(),
                   kls=[k.detach().clone() for k in kls], loss=loss.detach().clone())

    # ---- reference, eval mode (no power iteration), on the post-step buffers ---------------------
    model.eval()
    with torch.no_grad(), ref_import.patched_randn_like(eps):
        xe, rle, klse, msee = model(x)
    ref_eval = dict(x_hat=xe.clone(), recon=rle.clone(), mse=msee.clone(), kls=[k.clone() for k in klse])

    # ---- oracle restatement on the same inputs ----------------------------------------------------
    p = O.params_from_state_dict(sd0)
    ox, orl, okls, omse = O.vae_forward(p, x, eps, cfg["latent_dim"], cfg["lossfun"], training=True)
    oloss = O.total_loss(orl, okls, ALPHA, BETA)
    oloss.backward()
    worst = 0.0
    worst = max(worst, rel(ox, ref_out["x_hat"]), rel(orl, ref_out["recon"]), rel(omse, ref_out["mse"]))
    for a, b in zip(okls, ref_out["kls"]):
        worst = max(worst, rel(a, b))
    for n, gref in grads.items():
        if gref is None:
            assert p[n].grad is None, f"{name}: oracle produced a grad for dead parameter {n}"
        else:
            assert p[n].grad is not None, f"{name}: oracle misses grad for {n}"
            worst = max(worst, rel(p[n].grad, gref))
    for k in sd1:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            worst = max(worst, rel(p[k], sd1[k]))
    pe = {k: v for k, v in sd1.items()}
    with torch.no_grad():
        ex, erl, ekls, emse = O.vae_forward(pe, x, eps, cfg["latent_dim"], cfg["lossfun"], training=False)
    worst = max(worst, rel(ex, ref_eval["x_hat"]), rel(erl, ref_eval["recon"]))
    print(f"{name}: oracle vs reference worst rel-L2 = {worst:.3e}  "
          f"(params {sum(v.numel() for v in sd0.values())}, dead grads "
          f"{sum(1 for g_ in grads.values() if g_ is None)})")
    assert worst < 2e-5, f"{name}: oracle does not match the reference ({worst})"

    uv_after = {k: v for k, v in sd1.items() if k.endswith('weight_u') or k.endswith('weight_v')}
    torch.save(dict(cfg=cfg, alpha=ALPHA, beta=BETA, state_dict=sd0, uv_after=uv_after, x=x, eps=eps,
                    ref=ref_out, ref_eval=ref_eval, grads=grads), os.path.join(GOLDEN, name + ".pt"))


def run_elbo_curve(name="elbo_curve_toy3", steps=100):
    """100 optimiser steps through the reference modules with train.py's step semantics
    (train.py:139-168: zero_grad, forward, alpha*recon + beta*sum(kl), backward, AdamW.step)."""
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20, small=True,
               batch=8, lossfun="MSE")
    model = ref_import.build_reference_vae(cfg, seed=3)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B, N, T = cfg["batch"], cfg["num_node"], cfg["num_time"]
    data = O.synthetic_field(16, N, T, seed=21)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(9)
    epochs = 10                      # beta schedule: 10 "epochs" of 10 steps (2 batches x 5 repeats)
    curve, recon_curve, kl_curve, eps_all = [], [], [], []
    model.train(True)
    for step in range(steps):
        epoch = step // 10
        xb = data[(step % 2) * B:(step % 2) * B + B]
        eps = [torch.randn(s, generator=g) for s in O.eps_shapes(cfg, B)]
        eps_all.append(eps)
        opt.zero_grad(set_to_none=True)
        with ref_import.patched_randn_like(eps):
            _, rl, kls, _ = model(xb)
        beta = O.warmup_beta(epoch, epochs)
        loss = rl * ALPHA + sum(kls) * beta
        loss.backward()
        opt.step()
        curve.append(float(loss))
        recon_curve.append(float(rl))
        kl_curve.append(float(sum(kls)))
    print(f"{name}: loss {curve[0]:.4e} -> {curve[-1]:.4e}")
    torch.save(dict(cfg=cfg, alpha=ALPHA, epochs=epochs, lr=1e-3, state_dict=sd0, data=data, eps=eps_all,
                    loss=curve, recon=recon_curve, kl=kl_curve), os.path.join(GOLDEN, name + ".pt"))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    for n, c in CASES.items():
        run_case(n, c)
    run_elbo_curve()
