"""CPU tests of the Trainer's host logic (gradient sink / persistent arena, optimiser plan, data-parallel
all-reduce over gloo) with the CUDA kernels replaced by their torch models (tests/kernel_emulator.py)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from conftest import load_golden, rel_l2
from test_wiring_cpu import build_engine_vae


def _reference_steps(g, n_steps, xs):
    """torch autograd through the emulated engine + torch.optim.AdamW = the reference's step semantics."""
    m = build_engine_vae(g["cfg"], g["state_dict"])
    m.train(True)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    losses = []
    for i in range(n_steps):
        opt.zero_grad(set_to_none=True)
        with sg.fixed_eps(g["eps"]):
            _, rl, kls, _ = m(xs[i])
        loss = rl * g["alpha"] + sum(kls) * g["beta"]
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return m, losses


@pytest.mark.parametrize("name", ["toy3_small_mse", "toy4_large_mae"])
@pytest.mark.parametrize("fused", [True, False])
def test_trainer_matches_torch_adamw(name, fused):
    from simulgen_vae_b200.trainer import Trainer
    g = load_golden(name)
    sg.set_precision("fp32")
    try:
        with emu.install():
            xs = [g["x"], g["x"] * 0.9, g["x"] * 1.1]
            ref, ref_losses = _reference_steps(g, 3, xs)
            m = build_engine_vae(g["cfg"], g["state_dict"])
            m.train(True)
            tr = Trainer(m, lr=1e-3, alpha=g["alpha"], fused=fused)
            losses = []
            for i in range(3):
                with sg.fixed_eps(g["eps"]):
                    out = tr.step(xs[i], beta=g["beta"])
                losses.append(float(out[0]))
            if fused:
                assert all(p.grad is None for p in m.parameters())       # gradients live in the arena only
                assert tr.plan is not None and tr.sink.frozen
                n_live = sum(1 for v in g["grads"].values() if v is not None)
                assert tr.plan.n == n_live
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) / abs(b) < 1e-5
    rsd, sd = ref.state_dict(), m.state_dict()
    for k in rsd:
        assert rel_l2(sd[k], rsd[k]) < 2e-5, k
    assert tr.scalars()[4] > 0


def test_fp16_dynamic_loss_scale_equals_static_and_skips_overflow():
    """fp16 mode: the device-resident dynamic scaler (default) gives the weights of a static power-of-two scale while no
    overflow occurs; a forced overflow (absurd scale) skips the step - weights, moments and the applied-step counter
    unchanged - and halves the scale; training then continues."""
    from simulgen_vae_b200.trainer import Trainer
    g = load_golden("toy3_small_mse")
    sg.set_precision("fp16")
    try:
        with emu.install():
            def run(**kw):
                m = build_engine_vae(g["cfg"], g["state_dict"])
                m.train(True)
                tr = Trainer(m, lr=1e-3, alpha=g["alpha"], **kw)
                for i in range(3):
                    with sg.fixed_eps(g["eps"]):
                        tr.step(g["x"], beta=g["beta"])
                return m, tr
            m_dyn, tr_dyn = run()
            st = tr_dyn.scaler_state()
            assert st["step"] == 3 and st["skipped"] == 0 and st["good_steps"] == 3
            m_static, tr_static = run(loss_scale=st["scale"])
            assert tr_static.scaler is None
            sd_a, sd_b = m_dyn.state_dict(), m_static.state_dict()
            for k in sd_a:
                assert torch.equal(sd_a[k], sd_b[k]), k
            # overflow: fp16 gradients cannot hold a loss scaled by 2^60
            before = {k: v.clone() for k, v in sd_a.items() if not k.endswith(("weight_u", "weight_v"))}
            tr_dyn.scaler[:1].view(torch.float32).fill_(2.0 ** 60)
            with sg.fixed_eps(g["eps"]):
                tr_dyn.step(g["x"], beta=g["beta"])
            st2 = tr_dyn.scaler_state()
            assert st2["last_skipped"] == 1 and st2["skipped"] == 1 and st2["step"] == 3 and st2["scale"] == 2.0 ** 59
            assert tr_dyn.scalars()[4] == float("inf")
            after = m_dyn.state_dict()
            for k, v in before.items():
                assert torch.equal(after[k], v), k
            tr_dyn.scaler[:1].view(torch.float32).fill_(st["scale"])
            with sg.fixed_eps(g["eps"]):
                tr_dyn.step(g["x"], beta=g["beta"])
            st3 = tr_dyn.scaler_state()
            assert st3["last_skipped"] == 0 and st3["step"] == 4
            assert any(not torch.equal(m_dyn.state_dict()[k], v) for k, v in before.items())
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, name, fused, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    # exercise the chunked weight-gradient commit (k = 1 layers produced in row chunks, each all-reduced on its own) on
    # the toy model: every k = 1 layer with >= 16 output channels, 8-row chunks
    os.environ.update(SIMULGEN_B200_CHUNK_WGRAD_MELEMS="0", SIMULGEN_B200_CHUNK_WGRAD_ALIGN="8")
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from simulgen_vae_b200.trainer import Trainer
    g = load_golden(name)
    B = g["x"].shape[0]
    half = B // world
    sg.set_precision("fp32")
    with emu.install():
        m = build_engine_vae(g["cfg"], g["state_dict"])
        m.train(True)
        tr = Trainer(m, lr=1e-3, alpha=g["alpha"], fused=fused, bucket_mb=0)      # bucket_mb=0: one bucket per layer
        for i in range(2):
            eps = [e[rank * half:(rank + 1) * half] for e in g["eps"]]
            with sg.fixed_eps(eps):
                tr.step(g["x"][rank * half:(rank + 1) * half], beta=g["beta"], sample_offset=rank * half)
    if rank == 0:
        torch.save({k: v.clone() for k, v in m.state_dict().items()}, out)
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("fused", [True, False])
def test_data_parallel_world2_equals_single_process(tmp_path, fused):
    """Two gloo ranks, each with half of the batch, must reproduce the single-process full-batch weights:
    every op is per-sample and the losses are batch means."""
    from simulgen_vae_b200.trainer import Trainer
    name = "toy3_small_mse"
    g = load_golden(name)
    if g["x"].shape[0] % 2:
        pytest.skip("odd batch")
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), name, fused, out), nprocs=2, join=True)
    dp = torch.load(out)
    sg.set_precision("fp32")
    try:
        with emu.install():
            m = build_engine_vae(g["cfg"], g["state_dict"])
            m.train(True)
            tr = Trainer(m, lr=1e-3, alpha=g["alpha"], fused=True)
            for i in range(2):
                with sg.fixed_eps(g["eps"]):
                    tr.step(g["x"], beta=g["beta"])
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    sd = m.state_dict()
    for k in sd:
        assert rel_l2(dp[k], sd[k]) < 5e-5, (k, rel_l2(dp[k], sd[k]))


STATIC_CFG = dict(latent_dim=32, hierarchical_dim=8, enc=[64, 32], num_node=136, num_time=1, small=True, lossfun="MSE", batch=16)


def _static_batch_and_eps(B):
    from oracle import vae_oracle as O
    g = torch.Generator().manual_seed(12)
    x = torch.rand(B, 136, 1, generator=g) * 1.4 - 0.7
    eps = [torch.randn(s, generator=g) for s in O.eps_shapes(STATIC_CFG, B)]
    return x, eps


def _dp_static_worker(rank, world, port, sd_path, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.update(SIMULGEN_B200_CHUNK_WGRAD_MELEMS="0", SIMULGEN_B200_CHUNK_WGRAD_ALIGN="8")
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from simulgen_vae_b200.trainer import Trainer
    sg.set_precision("fp16")
    x, eps = _static_batch_and_eps(16)
    half = 16 // world
    with emu.install():
        m = build_engine_vae(STATIC_CFG, torch.load(sd_path))
        m.train(True)
        tr = Trainer(m, lr=1e-3, alpha=1e4, bucket_mb=0, loss_scale=1.0)
        for i in range(2):
            with sg.fixed_eps([e[rank * half:(rank + 1) * half] for e in eps]):
                tr.step(x[rank * half:(rank + 1) * half], beta=1e-2, sample_offset=rank * half)
    if rank == 0:
        torch.save({k: v.clone() for k, v in m.state_dict().items()}, out)
    torch.distributed.destroy_process_group()


def test_data_parallel_static_fields_compact_path(tmp_path):
    """Static fields (T = 1) in fp16 mode: two gloo ranks with 8 samples each run the compact [C][B] path of the two
    N-channel layers (per-rank batch a multiple of 8) and must reproduce the single-process step on all 16 samples."""
    from simulgen_vae_b200.trainer import Trainer
    sg.set_precision("fp16")
    try:
        with emu.install():
            torch.manual_seed(11)
            m = build_engine_vae(STATIC_CFG, None)
            sd_path = str(tmp_path / "init.pt")
            torch.save({k: v.clone() for k, v in m.state_dict().items()}, sd_path)
            out = str(tmp_path / "rank0.pt")
            mp.spawn(_dp_static_worker, args=(2, _free_port(), sd_path, out), nprocs=2, join=True)
            from simulgen_vae_b200 import kernels as K
            calls = {"n": 0}
            orig = K.static_recon_fwd

            def counting(*a, **k):
                calls["n"] += 1
                return orig(*a, **k)
            K.static_recon_fwd = counting
            m.train(True)
            x, eps = _static_batch_and_eps(16)
            tr = Trainer(m, lr=1e-3, alpha=1e4, loss_scale=1.0)
            for i in range(2):
                with sg.fixed_eps(eps):
                    tr.step(x, beta=1e-2)
            assert calls["n"] == 2
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    dp, sd = torch.load(out), m.state_dict()
    for k in sd:
        if k.endswith("weight_orig"):                      # matrices: averaged over many elements; vectors see Adam sign noise
            a, b = dp[k], sd[k]
            if a.dim() == 3 and a.shape[2] > 1:
                # at T = 1 only the centre tap of a k-tap conv has a data gradient; the other taps get the spectral-norm
                # correction alone (~1e-8: fp32 rounding noise that AdamW's m / sqrt(v) turns into +-lr steps, with or
                # without the centre-tap GEMMs) - they are compared on the centre tap
                a, b = a[:, :, a.shape[2] // 2], b[:, :, b.shape[2] // 2]
            assert rel_l2(a, b) < 5e-4, (k, rel_l2(a, b))


def _dp_unseeded_worker(rank, world, port, name, out_dir):
    """Like the reference's entry point: every rank draws its OWN initial weights and spectral-norm vectors."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from simulgen_vae_b200.trainer import Trainer
    g = load_golden(name)
    half = g["x"].shape[0] // world
    sg.set_precision("fp32")
    torch.manual_seed(1000 + 17 * rank)                                # different replicas before the broadcast
    with emu.install():
        m = build_engine_vae(g["cfg"], None)
        m.train(True)
        before = {k: v.clone() for k, v in m.state_dict().items()}
        tr = Trainer(m, lr=1e-3, alpha=g["alpha"], bucket_mb=0)
        after_init = {k: v.clone() for k, v in m.state_dict().items()}
        for i in range(2):
            eps = [e[rank * half:(rank + 1) * half] for e in g["eps"]]
            with sg.fixed_eps(eps):
                tr.step(g["x"][rank * half:(rank + 1) * half], beta=g["beta"], sample_offset=rank * half)
    torch.save(dict(before=before, init=after_init, final={k: v.clone() for k, v in m.state_dict().items()}),
               os.path.join(out_dir, "rank%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_data_parallel_replicas_are_broadcast_from_rank0(tmp_path):
    """ranks that build their model without a common seed must still train ONE model - Trainer
    broadcasts rank 0's parameters and buffers (incl. weight_u / weight_v) before the first step."""
    name = "toy3_small_mse"
    mp.spawn(_dp_unseeded_worker, args=(2, _free_port(), name, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(str(tmp_path / "rank0.pt"))
    r1 = torch.load(str(tmp_path / "rank1.pt"))
    differ = [k for k in r0["before"] if not torch.equal(r0["before"][k], r1["before"][k])]
    assert any(k.endswith("weight_orig") for k in differ) and any(k.endswith("weight_u") for k in differ)
    for k in r0["init"]:
        assert torch.equal(r0["init"][k], r0["before"][k]), k          # rank 0 is the source
        assert torch.equal(r1["init"][k], r0["init"][k]), k
        assert torch.equal(r1["final"][k], r0["final"][k]), k          # and the replicas stay identical
    assert any(not torch.equal(r0["final"][k], r0["init"][k]) for k in r0["init"])


def _train_loop_worker(rank, world, port, workdir, data):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    if world > 1:
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    os.makedirs(os.path.join(workdir, "w%d_r%d" % (world, rank)), exist_ok=True)
    os.chdir(os.path.join(workdir, "w%d_r%d" % (world, rank)))
    torch.randn_like = lambda t, *a, **k: torch.zeros_like(t)          # reparameterisation noise off: DP == single exactly
    from simulgen_vae_b200 import train_loop
    sg.install_overlay()
    sg.set_precision("fp32")
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)
    B = 4
    per = B // world
    # every global batch of 4 is split rank-major: rank r takes samples [r*per, (r+1)*per) of it
    idx = [b * B + rank * per + j for b in range(3) for j in range(per)]
    train_dl = torch.utils.data.DataLoader(data[idx], batch_size=per, shuffle=False)
    vidx = [12 + rank * per + j for j in range(per)]
    val_dl = torch.utils.data.DataLoader(data[vidx], batch_size=per, shuffle=False)
    torch.manual_seed(3)
    with emu.install():
        train_loop.train(4, B, train_dl, val_dl, 1e-3, cfg["enc"], cfg["enc"][::-1], cfg["num_node"], cfg["latent_dim"],
                         cfg["hierarchical_dim"], cfg["num_time"], 1000000, "MSE", True, True, device="cpu")
    if world > 1:
        torch.distributed.destroy_process_group()


def test_train_driver_data_parallel_world2_equals_single_process(tmp_path):
    """simulgen_vae_b200.train_loop.train under torch.distributed (gloo, 2 ranks, half of every batch each) ends with the
    weights of the single-process run on the full batches; only rank 0 writes the checkpoints."""
    g = torch.Generator().manual_seed(0)
    data = torch.rand(16, 64, 20, generator=g) * 1.4 - 0.7
    mp.spawn(_train_loop_worker, args=(2, _free_port(), str(tmp_path), data), nprocs=2, join=True)
    mp.spawn(_train_loop_worker, args=(1, _free_port(), str(tmp_path), data), nprocs=1, join=True)
    dp = torch.load(str(tmp_path / "w2_r0" / "checkpoints" / "SimulGen-VAE.pth"), weights_only=False)
    single = torch.load(str(tmp_path / "w1_r0" / "checkpoints" / "SimulGen-VAE.pth"), weights_only=False)
    assert not os.path.exists(str(tmp_path / "w2_r1" / "checkpoints" / "SimulGen-VAE.pth"))
    assert list(dp.keys()) == list(single.keys())
    for k in single:
        assert rel_l2(dp[k], single[k]) < 1e-4, (k, rel_l2(dp[k], single[k]))


def _loader_dp_worker(rank, world, port, workdir, data):
    import random
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    if world > 1:
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    os.makedirs(os.path.join(workdir, "w%d_r%d" % (world, rank)), exist_ok=True)
    os.chdir(os.path.join(workdir, "w%d_r%d" % (world, rank)))
    torch.randn_like = lambda t, *a, **k: torch.zeros_like(t)          # reparameterisation noise off: DP == single exactly
    # like the reference's entry point nobody seeds the ranks alike: rank 0 == the single process, rank 1 differs
    random.seed(5 + 100 * rank); np.random.seed(6 + 100 * rank); torch.manual_seed(7 + 100 * rank)
    sg.install_overlay(train=True)
    sg.set_precision("fp32")
    from modules.augmentation import create_augmented_dataloaders      # the overlay's drop-in
    from modules.train import train
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16, 8], num_node=64, num_time=20)
    batch_size = 4 // world                                            # SimulGen-VAE.py:172: Batch_size //= world_size
    with emu.install():
        train_dl, val_dl = create_augmented_dataloaders(data.numpy(), batch_size, load_all=True, device="cpu")
        for dl in (train_dl, val_dl):
            dl.injected_noise = lambda levels, shape: torch.zeros((len(levels),) + tuple(shape))
        assert len(train_dl) == 4 and len(val_dl) == 1
        curves = train(4, batch_size, train_dl, val_dl, 1e-3, cfg["enc"], cfg["enc"][::-1], cfg["num_node"], cfg["latent_dim"],
                       cfg["hierarchical_dim"], cfg["num_time"], 1000000, "MSE", True, True, device="cpu")
    torch.save([torch.as_tensor(c) for c in curves], "curves.pt")
    if world > 1:
        torch.distributed.destroy_process_group()


def test_overlay_loader_shards_global_batches_across_ranks(tmp_path):
    """`SimulGen-VAE.py --use_ddp` + install_overlay(train=True) without hand-sharding: the overlay's
    create_augmented_dataloaders hands rank r its 1/W slice of every global batch (one split, one shuffle, one set of
    augmentation decisions - rank 0's, broadcast), so 2 gloo ranks at per-rank batch 2 end with the weights and the loss
    curves of the single process at batch 4."""
    g = torch.Generator().manual_seed(0)
    data = torch.rand(20, 64, 20, generator=g) * 1.4 - 0.7
    mp.spawn(_loader_dp_worker, args=(2, _free_port(), str(tmp_path), data), nprocs=2, join=True)
    mp.spawn(_loader_dp_worker, args=(1, _free_port(), str(tmp_path), data), nprocs=1, join=True)
    dp = torch.load(str(tmp_path / "w2_r0" / "checkpoints" / "SimulGen-VAE.pth"), weights_only=False)
    single = torch.load(str(tmp_path / "w1_r0" / "checkpoints" / "SimulGen-VAE.pth"), weights_only=False)
    for k in single:
        assert rel_l2(dp[k], single[k]) < 1e-4, (k, rel_l2(dp[k], single[k]))
    c_dp = torch.load(str(tmp_path / "w2_r0" / "curves.pt"))
    c_1 = torch.load(str(tmp_path / "w1_r0" / "curves.pt"))
    for a, b in zip(c_dp, c_1):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), (a, b)
