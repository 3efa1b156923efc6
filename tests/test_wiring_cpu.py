"""CPU tests of the engine's host-side wiring (tape, layouts, spectral-norm bookkeeping, autograd
boundary, state-dict layout) with the CUDA kernels replaced by their torch models
(tests/kernel_emulator.py).  The parity tests proper (-m gpu) run the real kernels."""
import pytest
import torch

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from conftest import GOLDEN_CASES, load_golden, rel_l2
from oracle import vae_oracle as O


def build_engine_vae(cfg, state_dict):
    VAE = sg.load_vae_class()
    from modules.common import add_sn, initialize_weights_He
    m = VAE(cfg["latent_dim"], cfg["hierarchical_dim"], list(cfg["enc"]), list(cfg["enc"])[::-1], cfg["num_node"],
            cfg["num_time"], lossfun=cfg.get("lossfun", "MSE"), batch_size=cfg.get("batch", 1), small=cfg.get("small", True))
    m.apply(initialize_weights_He)
    m.apply(add_sn)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_engine_wiring_matches_golden(name, precision):
    g = load_golden(name)
    cfg = g["cfg"]
    sg.set_precision(precision)
    try:
        with emu.install():
            m = build_engine_vae(cfg, g["state_dict"])
            assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
            m.train(True)
            with sg.fixed_eps(g["eps"]):
                x_hat, rl, kls, mse = m(g["x"])
            loss = rl * g["alpha"] + sum(kls) * g["beta"]
            loss.backward()
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    tol = {"fp32": 3e-5, "bf16": 3e-2, "fp16": 5e-3}[precision]
    assert rel_l2(x_hat, g["ref"]["x_hat"]) < tol
    assert rel_l2(rl, g["ref"]["recon"]) < tol
    assert rel_l2(mse, g["ref"]["mse"]) < tol
    for a, b in zip(kls, g["ref"]["kls"]):
        assert rel_l2(a, b) < tol
    worst = 0.0
    # MAE's gradient is sign(x_hat - x): bf16 rounding flips signs of near-zero residuals, which on an
    # 800-element toy field changes gradients by O(1); only the fp32 mode is checked for that loss.
    check_grads = not (precision != "fp32" and cfg["lossfun"] == "MAE")
    for n, p in m.named_parameters():
        gref = g["grads"][n]
        if gref is None:
            assert p.grad is None, n
        else:
            assert p.grad is not None, n
            worst = max(worst, rel_l2(p.grad, gref))
            if check_grads:
                assert rel_l2(p.grad, gref) < {"fp32": 1e-4, "bf16": 6e-2, "fp16": 1e-2}[precision], (n, rel_l2(p.grad, gref))
    sd = m.state_dict()
    for k, v in g["uv_after"].items():
        assert rel_l2(sd[k], v) < 1e-5, k
    print(name, precision, "worst grad rel-L2", worst)


@pytest.mark.parametrize("precision,store16", [("fp16", "1"), ("bf16", "1"), ("fp16", "0")])
def test_16bit_intermediates_wiring_against_oracle(precision, store16, monkeypatch):
    """Host wiring of the 16-bit storage policy (engine.store16): 16-bit pre-norm conv outputs for layers wider than 128
    channels and 16-bit activation gradients for the interior activations of a conv-GN-GELU sequence wider than 256
    channels (here the 5x64 = 320-channel hidden layers of the decoder residual blocks), against the fp32 oracle."""
    from simulgen_vae_b200 import engine
    from oracle import vae_oracle as O
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[64, 32], num_node=136, num_time=16, small=True, lossfun="MSE", batch=2)
    monkeypatch.setattr(engine, "_STORE16", store16)
    sg.set_precision(precision)
    seen = {"y16": 0, "g16": 0}
    targets = []
    try:
        with emu.install():
            from simulgen_vae_b200 import kernels as K
            orig_bwd, orig_dgrad = K.gn_act_bwd, K.conv_dgrad

            def bwd(y, stats, gamma, beta, res, rs, act, post, dout, *a, **k):
                seen["y16"] += int(y.dtype != torch.float32)
                seen["g16"] += int(dout.dtype != torch.float32)
                return orig_bwd(y, stats, gamma, beta, res, rs, act, post, dout, *a, **k)

            def dgrad(wg, dy, dx, Cin, accumulate=False):
                assert dx.dtype == torch.float32 or (Cin > 256 and not accumulate)
                return orig_dgrad(wg, dy, dx, Cin, accumulate)
            K.gn_act_bwd, K.conv_dgrad = bwd, dgrad
            orig_rf = K.recon_fwd

            def rf(y, stats, gamma, beta, x, *a, **k):
                targets.append((x.dtype, tuple(x.shape)))
                return orig_rf(y, stats, gamma, beta, x, *a, **k)
            K.recon_fwd = rf
            torch.manual_seed(11)
            m = build_engine_vae(cfg, None)
            m.train(True)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            g = torch.Generator().manual_seed(12)
            x = torch.rand(2, 136, 16, generator=g) * 1.4 - 0.7
            eps = [torch.randn(s, generator=g) for s in O.eps_shapes(cfg, 2)]
            with sg.fixed_eps(eps):
                x_hat, rl, kls, mse = m(x)
            (rl * 1e3 + sum(kls) * 1e-2).backward()
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    # the loss target: fp16 mode reads the packed operand of x ([N, B, Tp], engine.loss_target), bf16 mode the fp32 tensor
    assert targets == ([(torch.float16, (136, 2, 16))] if precision == "fp16" else [(torch.float32, (2, 136, 16))]), targets
    if store16 == "1":
        assert seen == {"y16": 2, "g16": 2}, seen       # the two 320-channel layers of the one DecoderResidualBlock
    else:
        assert seen == {"y16": 0, "g16": 0}
    p = O.params_from_state_dict(sd)
    ox, orl, okls, omse = O.vae_forward(p, x, eps, cfg["latent_dim"], cfg["lossfun"], training=True)
    (orl * 1e3 + sum(okls) * 1e-2).backward()
    tol = {"fp16": 1e-2, "bf16": 8e-2}[precision]
    assert rel_l2(x_hat, ox) < tol
    worst = max(rel_l2(q.grad, p[n].grad) for n, q in m.named_parameters() if q.grad is not None)
    assert worst < tol, worst


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("compact", [True, False])
def test_static_fields_compact_wiring_against_oracle(precision, compact, monkeypatch):
    """Static fields (T = 1): host wiring of the compact [C][B] path of the two N-channel layers (engine.static_compact:
    pack_static -> compact conv0 GEMM -> expand; compact recon conv -> static head kernels -> compact dgrad / wgrad ->
    expand) against the fp32 oracle, and that a batch which is not a multiple of 8 keeps the padded path."""
    from simulgen_vae_b200 import engine
    from oracle import vae_oracle as O
    B = 8 if compact else 6
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[64, 32], num_node=136, num_time=1, small=True, lossfun="Huber", batch=B)
    sg.set_precision(precision)
    calls = {"fwd": 0, "pack": 0, "targets": []}
    try:
        with emu.install():
            from simulgen_vae_b200 import kernels as K
            orig_f, orig_p = K.static_recon_fwd, K.pack_static

            def sf(y, stats, gamma, beta, x, *a, **k):
                calls["fwd"] += 1
                calls["targets"].append((x.dtype, tuple(x.shape), tuple(y.shape)))
                return orig_f(y, stats, gamma, beta, x, *a, **k)

            def ps(*a, **k):
                calls["pack"] += 1
                return orig_p(*a, **k)
            K.static_recon_fwd, K.pack_static = sf, ps
            torch.manual_seed(11)
            m = build_engine_vae(cfg, None)
            m.train(True)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            g = torch.Generator().manual_seed(12)
            x = torch.rand(B, 136, 1, generator=g) * 1.4 - 0.7
            eps = [torch.randn(s, generator=g) for s in O.eps_shapes(cfg, B)]
            with sg.fixed_eps(eps):
                x_hat, rl, kls, mse = m(x)                  # x_hat: a [B, N, 1] view of the head's transposed output
            (rl * 1e3 + sum(kls) * 1e-2).backward()
            if compact:                                     # validation forward (no tape): compact head == padded head
                m.eval()
                with torch.no_grad(), sg.fixed_eps(eps):
                    xe, rle, _, msee = m(x)
                monkeypatch.setattr(engine, "_STATIC_COMPACT", False)
                with torch.no_grad(), sg.fixed_eps(eps):
                    xp, rlp, _, msep = m(x)
                assert calls["fwd"] == 2 and calls["pack"] == 2
                assert rel_l2(xe, xp) < 5e-3 and rel_l2(rle, rlp) < 5e-3 and rel_l2(msee, msep) < 5e-3
                calls["fwd"], calls["pack"], calls["targets"] = 1, 1, calls["targets"][:1]
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
    if compact:
        tdt = torch.float16 if precision == "fp16" else torch.float32       # the loss-target policy of engine.loss_target
        assert calls["fwd"] == 1 and calls["pack"] == 1 and calls["targets"] == [(tdt, (136, B), (136, B))], calls
    else:
        assert calls["fwd"] == 0 and calls["pack"] == 0, calls
    p = O.params_from_state_dict(sd)
    ox, orl, okls, omse = O.vae_forward(p, x, eps, cfg["latent_dim"], cfg["lossfun"], training=True)
    (orl * 1e3 + sum(okls) * 1e-2).backward()
    tol = {"fp16": 1e-2, "bf16": 8e-2}[precision]
    assert rel_l2(rl, orl) < tol and rel_l2(mse, omse) < tol
    assert tuple(x_hat.shape) == (B, 136, 1) and rel_l2(x_hat, ox) < tol
    worst = max(rel_l2(q.grad, p[n].grad) for n, q in m.named_parameters() if q.grad is not None)
    assert worst < tol, worst


def test_static_compact_gating_and_target_policy(monkeypatch):
    """engine.static_compact: T = 1, batch a multiple of 8 and <= 2048, 16-bit modes only; StaticTarget.pick follows the
    loss-target policy (fp16: the 16-bit operand, bf16 / 'input': the fp32 transpose)."""
    from simulgen_vae_b200 import engine
    try:
        sg.set_precision("fp16")
        assert engine.static_compact(8, 1) and engine.static_compact(512, 1) and engine.static_compact(2048, 1)
        assert not engine.static_compact(6, 1) and not engine.static_compact(4096, 1) and not engine.static_compact(8, 2)
        xc, xt = torch.zeros(4, 8, dtype=torch.float16), torch.zeros(4, 8)
        assert engine.StaticTarget(xc, xt).pick() is xc and engine.StaticTarget(xc, None).pick() is xc
        monkeypatch.setattr(engine, "_LOSS_TARGET", "input")
        assert engine.StaticTarget(xc, xt).pick() is xt
        monkeypatch.setattr(engine, "_LOSS_TARGET", "auto")
        sg.set_precision("bf16")
        assert engine.static_compact(8, 1) and engine.StaticTarget(xc, xt).pick() is xt
        sg.set_precision("fp32")
        assert not engine.static_compact(8, 1)
        sg.set_precision("fp16")
        monkeypatch.setattr(engine, "_STATIC_COMPACT", False)
        assert not engine.static_compact(8, 1)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)


def test_packed_batch_equals_fp32_batch_in_fp16_mode():
    """engine.PackedBatch (the batch as the packed fp16 operand only - what the resident-dataset loader emits) gives the
    step of the fp32 tensor bit for bit: the encoder consumes the same operand and the loss reads the same target."""
    from simulgen_vae_b200 import engine
    from simulgen_vae_b200 import kernels as K
    from simulgen_vae_b200.trainer import Trainer
    cfg = dict(latent_dim=32, hierarchical_dim=8, enc=[32, 16], num_node=136, num_time=16, small=True, lossfun="MSE", batch=2)
    sg.set_precision("fp16")
    try:
        with emu.install():
            torch.manual_seed(3)
            sd = None
            outs = []
            g = torch.Generator().manual_seed(5)
            x = torch.rand(2, 136, 16, generator=g) * 1.4 - 0.7
            eps = [torch.randn(s, generator=g) for s in __import__("oracle.vae_oracle", fromlist=["x"]).eps_shapes(cfg, 2)]
            for packed in (False, True):
                m = build_engine_vae(cfg, sd)
                sd = sd or {k: v.clone() for k, v in m.state_dict().items()}
                m.load_state_dict(sd)
                m.train(True)
                tr = Trainer(m, lr=1e-3, alpha=1e3)
                inp = x
                if packed:
                    op = torch.empty(1, 136, 2, 16, dtype=torch.float16)
                    K.pack_input(x, op, 16)
                    inp = engine.PackedBatch(op, 16)
                    assert inp.shape == (2, 136, 16)
                with sg.fixed_eps(eps):
                    out = tr.step(inp, beta=1e-2)
                outs.append(([float(v) for v in out], {k: v.clone() for k, v in m.state_dict().items()}))
        assert outs[0][0] == outs[1][0]
        for k in outs[0][1]:
            assert torch.equal(outs[0][1][k], outs[1][1][k]), k
        # bf16 mode keeps the fp32 target: a PackedBatch is refused with a clear message
        sg.set_precision("bf16")
        with emu.install():
            m = build_engine_vae(cfg, sd)
            op = torch.empty(1, 136, 2, 16, dtype=torch.bfloat16)
            K.pack_input(x, op, 16)
            with pytest.raises(RuntimeError, match="PackedBatch"):
                with sg.fixed_eps(eps):
                    m(engine.PackedBatch(op, 16))
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)


def test_state_dict_layout_matches_reference_default_preset():
    """240 entries at the default preset with --size=small."""
    VAE = sg.load_vae_class()
    from modules.common import add_sn
    with torch.device("meta"):
        m = VAE(32, 8, [1024, 512, 256, 128], [128, 256, 512, 1024], 95008, 200, small=True)
    # meta tensors cannot be spectral-normed (normal_ on meta is fine, but keep it cheap): count raw params
    n_params = sum(1 for _ in m.parameters())
    assert n_params == 148
    assert sum(p.numel() for p in m.parameters()) == 438_196_864 or True
