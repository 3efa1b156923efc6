"""-m gpu: every streaming / small CUDA kernel against its plain PyTorch model (tests/kernel_emulator.py),
called through the C ABI (simulgen_vae_b200.kernels -> libsimulgen_b200.so)."""
import pytest
import torch

import kernel_emulator as emu
from conftest import rel_l2
from simulgen_vae_b200 import kernels as K
from simulgen_vae_b200 import engine


def tp_of(T):
    """fp32-mode row pitch (zero gap of >= 2 columns): required by the SIMT convolutions, valid for every kernel"""
    return engine.tp_of(T, "fp32")

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def cr(C, B, T, seed=0, dtype=torch.float32, scale=1.0):
    """random CR-layout tensor with a zero gap"""
    Tp = tp_of(T)
    t = rnd(C, B, Tp, seed=seed, scale=scale)
    t[:, :, T:] = 0
    return t.to(dtype)


def planes_of(val, P, T, dtype=torch.float32):
    """operand tensor [P, C, B, Tp] holding `val` [C, B, Tp] in pre-shifted planes"""
    C, B, Tp = val.shape
    out = torch.empty(P, C, B, Tp, device=DEV, dtype=dtype)
    emu.write_planes(out, val[:, :, :T], T)
    return out


def close(a, b, tol, what=""):
    e = rel_l2(a, b)
    assert e < tol, "%s rel-L2 %.3e >= %.1e" % (what, e, tol)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,N,T", [(3, 37, 20), (70, 75, 1), (33, 64, 5), (3, 37, 8)])   # Tp == 8: short-row kernel (32 x 32 tiles)
def test_pack_unpack(dtype, B, N, T):
    x = rnd(B, N, T)
    out = torch.full((1, N, B, tp_of(T)), 7.0, device=DEV, dtype=dtype)
    ref = torch.empty_like(out)
    K.pack_input(x, out, T)
    emu.pack_input(x, ref, T)
    assert torch.equal(out, ref)
    if dtype == torch.float32:
        back = torch.empty(B, N, T, device=DEV)
        K.unpack_f32(out[0], back, T)
        assert torch.equal(back, x)


@pytest.mark.parametrize("kind", ["conv", "convT", "linear"])
@pytest.mark.parametrize("training", [True, False])
def test_sn_power_iter_pack_grad(kind, training):
    Cout, Cin, k = 40, 24, 3
    if kind == "conv":
        w = rnd(Cout, Cin, k)
        so, si, flip = Cin * k, k, 0
    elif kind == "convT":
        w = rnd(Cin, Cout, k)
        so, si, flip = k, Cout * k, 1
    else:
        k = 1
        w = rnd(Cout, Cin)
        so, si, flip = Cin, 1, 0
    u = torch.nn.functional.normalize(rnd(Cout, seed=1), dim=0)
    v = torch.nn.functional.normalize(rnd(Cin * k, seed=2), dim=0)
    u2, v2 = u.clone(), v.clone()
    s1, s2 = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    K.sn_power_iter(w, u, v, s1, Cout, Cin, k, so, si, training)
    emu.sn_power_iter(w, u2, v2, s2, Cout, Cin, k, so, si, training)
    close(u, u2, 1e-5, "u")
    close(v, v2, 1e-5, "v")
    close(s1, s2, 1e-5, "sigma")
    if kind == "linear":
        Cin_p = Cin
    else:
        Cin_p = (Cin + 7) // 8 * 8
        for dt in (torch.float32, torch.bfloat16):
            wg = torch.full((k, Cout, Cin_p), 3.0, device=DEV, dtype=dt)
            wg2 = torch.empty_like(wg)
            K.sn_pack_weight(w, s1, wg, Cout, Cin, Cin_p, k, so, si, flip)
            emu.sn_pack_weight(w, s1, wg2, Cout, Cin, Cin_p, k, so, si, flip)
            close(wg.float(), wg2.float(), 1e-6 if dt == torch.float32 else 4e-3, "wg")
    dwg = rnd(k, Cout, Cin_p, seed=5)
    g1, g2 = torch.empty_like(w), torch.empty_like(w)
    K.sn_weight_grad(dwg, w, u, v, s1, g1, Cout, Cin, Cin_p, k, so, si, flip)
    emu.sn_weight_grad(dwg, w, u, v, s1, g2, Cout, Cin, Cin_p, k, so, si, flip)
    close(g1, g2, 1e-5, "sn grad")


@pytest.mark.parametrize("k", [1, 3, 5])
def test_conv_simt_fp32(k):
    Cin, Cout, B, T = 20, 72, 3, 21
    Cin_p = (Cin + 7) // 8 * 8
    Tp = tp_of(T)
    wg = rnd(k, Cout, Cin_p, seed=1, scale=0.2)
    wg[:, :, Cin:] = 0
    act = planes_of(cr(Cin, B, T, seed=2), k + 2 if k < 5 else 5, T)
    bias = rnd(Cout, seed=3)
    o1 = torch.empty(Cout, B, Tp, device=DEV)
    o2 = torch.empty_like(o1)
    # gap columns t >= T are don't-care (the SIMT kernels convolve along the flattened axis, the model per sample)
    K.conv_fprop(wg, act, bias, o1, Cin)
    emu.conv_fprop(wg, act, bias, o2, Cin)
    close(o1[:, :, :T], o2[:, :, :T], 1e-5, "fprop")
    K.conv_fprop(wg, act, None, o1, Cin, accumulate=True)
    emu.conv_fprop(wg, act, None, o2, Cin, accumulate=True)
    close(o1[:, :, :T], o2[:, :, :T], 1e-5, "fprop acc")
    dy = planes_of(cr(Cout, B, T, seed=4), k, T)
    d1 = torch.empty(Cin, B, Tp, device=DEV)
    d2 = torch.empty_like(d1)
    K.conv_dgrad(wg, dy, d1, Cin)
    emu.conv_dgrad(wg, dy, d2, Cin)
    close(d1[:, :, :T], d2[:, :, :T], 1e-5, "dgrad")
    w1 = torch.empty(k, Cout, Cin_p, device=DEV)
    w2 = torch.empty_like(w1)
    K.conv_wgrad(dy, act, w1, Cin)
    emu.conv_wgrad(dy, act, w2, Cin)
    close(w1[:, :, :Cin], w2[:, :, :Cin], 1e-5, "wgrad")


def test_gn_stats():
    C, B, T, G = 48, 3, 21, 8
    y = cr(C, B, T, seed=1) + 0.5
    s1 = torch.empty(B, G, 2, device=DEV)
    s2 = torch.empty_like(s1)
    K.gn_stats(y, s1, T, G)
    emu.gn_stats(y, s2, T, G)
    close(s1, s2, 1e-6, "stats (mean, rstd)")


CASES = [
    # use_gn, act, res ('none'|'f32'|'op'), res_scale, post_gelu
    (True, K.ACT_GELU, "none", 1.0, False),
    (True, K.ACT_GELU, "f32", 0.1, False),
    (True, K.ACT_GELU, "op", 0.1, True),
    (False, K.ACT_GELU, "none", 1.0, False),
    (False, K.ACT_NONE, "none", 1.0, False),
    (True, K.ACT_TANH, "none", 1.0, False),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype,y16,d16", [(torch.bfloat16, False, False), (torch.float32, False, False),
                                           (torch.float16, False, False), (torch.float16, True, True),
                                           (torch.float16, True, False), (torch.bfloat16, False, True)])
@pytest.mark.parametrize("P,T", [(1, 21), (3, 21), (5, 21), (3, 300), (5, 200), (1, 200),
                                 (1, 1), (3, 1), (5, 1), (5, 5), (3, 8)])          # Tp == 8: short-row kernels
def test_gn_act_fwd_bwd(case, dtype, y16, d16, P, T):
    """T=21 / 200: rows of <= 256 elements (one segment per lane, shifted planes by warp shuffle);
    T=300: longer rows (several segments per lane, planes staged through shared memory).
    T<=8: short rows (static fields): one thread per row.
    y16 / d16: the pre-norm conv output / the incoming gradient stored in the 16-bit operand format (GroupNorm layers)."""
    use_gn, act, res_kind, res_scale, post = case
    if (y16 or d16) and not use_gn:
        pytest.skip("16-bit y / dout exist for GroupNorm layers only")
    C, B, G = 48, 19, 8                         # 19 samples: two (channel, 16-sample chunk) tasks, the second ragged
    if T <= 8:
        C, B = 80, 70                           # short rows: (8 + 2)-channel slices per group, 32-sample chunks, the last ragged
    Tp = tp_of(T)
    y = cr(C, B, T, seed=1) * 1.5 + 0.3
    if y16:
        y = y.to(dtype)
    gamma, beta = rnd(C, seed=2) * 0.5 + 1.0, rnd(C, seed=3) * 0.2
    stats = None
    if use_gn:
        stats = torch.empty(B, G, 2, device=DEV)
        K.gn_stats(y.float(), stats, T, G)
    res = None
    if res_kind == "f32":
        res = cr(C, B, T, seed=4)
    elif res_kind == "op":
        res = cr(C, B, T, seed=4).to(dtype)
    o1 = torch.full((P, C, B, Tp), 9.0, device=DEV, dtype=dtype)
    f1 = torch.full((C, B, Tp), 9.0, device=DEV)
    o2, f2 = torch.empty_like(o1), torch.empty_like(f1)
    GG = G if use_gn else 0
    K.gn_act_fwd(y, stats, gamma if use_gn else None, beta if use_gn else None, res, res_scale, act, post, o1, f1, T, GG)
    emu.gn_act_fwd(y, stats, gamma, beta, res, res_scale, act, post, o2, f2, T, G)
    close(f1, f2, 2e-6, "fwd f32")
    tol16 = {torch.float32: 2e-6, torch.bfloat16: 4e-3, torch.float16: 5e-4}[dtype]
    close(o1.float(), o2.float(), tol16, "fwd op")
    assert float(o1[:, :, :, T:].float().abs().max()) == 0.0
    # backward (GroupNorm layers overwrite dout: each side gets its own copy)
    dout = cr(C, B, T, seed=5)
    if d16:
        dout = dout.to(dtype)
    dout_k = dout.clone()
    dy1 = torch.full((P, C, B, Tp), 9.0, device=DEV, dtype=dtype)
    dy2 = torch.empty_like(dy1)
    dg1, db1, dbi1 = (torch.empty(C, device=DEV) for _ in range(3))
    dg2, db2, dbi2 = (torch.empty(C, device=DEV) for _ in range(3))
    dr1 = cr(C, B, T, seed=6) if res is not None else None
    dr2 = dr1.clone() if dr1 is not None else None
    K.gn_act_bwd(y, stats, gamma if use_gn else None, beta if use_gn else None, res, res_scale, act, post, dout_k, dy1,
                 dg1 if use_gn else None, db1 if use_gn else None, dbi1, dr1, 1, T, GG)
    emu.gn_act_bwd(y, stats, gamma, beta, res, res_scale, act, post, dout, dy2, dg2, db2, dbi2, dr2, 1, T, G)
    tol = {torch.float32: 2e-5, torch.bfloat16: 5e-3, torch.float16: 6e-4}[dtype]
    if d16:
        tol *= 1.5                               # dz makes one more 16-bit round trip between the two passes
    close(dy1.float(), dy2.float(), tol, "dy")
    close(dbi1, dbi2, tol * 2, "dbias")
    if use_gn:
        close(dg1, dg2, 3e-5, "dgamma")
        close(db1, db2, 3e-5, "dbeta")
    if res is not None:
        close(dr1[:, :, :T], dr2[:, :, :T], 2e-5, "dres")
    assert float(dy1[:, :, :, T:].float().abs().max()) == 0.0


@pytest.mark.parametrize("loss", ["MSE", "MAE", "smoothL1", "Huber"])
@pytest.mark.parametrize("with_ext,one_pass", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("T,y_bf16", [(21, False), (24, False), (24, True), (21, True),    # 24: fast path (T % 8 == 0)
                                      (1, False), (1, True), (5, True), (8, False)])         # Tp == 8: short-row kernels
def test_recon_fwd_bwd(loss, with_ext, one_pass, T, y_bf16):
    N, B, G = (40, 3, 8) if T > 8 else (72, 70, 8)         # short rows: 32 x 32 tiles with ragged edges on both axes
    Tp = tp_of(T)
    kind = K.LOSS_KINDS[loss]
    y = cr(N, B, T, seed=1) * 2.0
    gamma, beta = rnd(N, seed=2) * 0.5 + 1.0, rnd(N, seed=3) * 0.2
    x = rnd(B, N, T, seed=4) * 1.5
    stats = torch.empty(B, G, 2, device=DEV)
    K.gn_stats(y, stats, T, G)
    if y_bf16:                       # bf16-stored pre-norm output; the model recomputes the statistics from what it is given
        y = y.to(torch.bfloat16)
        K.gn_stats(y.float(), stats, T, G)
    xh1, xh2 = torch.empty(B, N, T, device=DEV), torch.empty(B, N, T, device=DEV)
    s1, s2 = (torch.empty(2, device=DEV, dtype=torch.float64) for _ in range(2))
    rowsums = torch.full((N * B, 4), 9.0, device=DEV) if one_pass else None
    K.recon_fwd(y, stats, gamma, beta, x, xh1, s1, T, G, kind, rowsums)
    emu.recon_fwd(y, stats, gamma, beta, x, xh2, s2, T, G, kind)
    close(xh1, xh2, 3e-6, "x_hat")
    close(s1, s2, 1e-5, "loss sums")
    g_loss = torch.tensor([1.0e6], device=DEV)
    g_mse = torch.tensor([0.5], device=DEV)
    ext = rnd(B, N, T, seed=7) * 0.01 if with_ext else None
    inv = 1.0 / (B * N * T)
    outs = []
    for fn in (K.recon_bwd, emu.recon_bwd):
        dy = torch.full((1, N, B, Tp), 9.0, device=DEV, dtype=torch.bfloat16 if y_bf16 else torch.float32)
        dg, db, dbi = (torch.empty(N, device=DEV) for _ in range(3))
        fn(y, stats, gamma, beta, x, g_loss, g_mse, inv, ext, dy, dg, db, dbi, T, G, kind, rowsums)
        outs.append((dy.float(), dg, db, dbi))
    for a, b, nm in zip(outs[0], outs[1], ("dy", "dgamma", "dbeta", "dbias")):
        close(a, b, 5e-3 if (y_bf16 and nm in ("dy", "dbias")) else 5e-5, nm)
    # ext-only path (decoder used without the fused loss)
    if with_ext:
        outs = []
        for fn in (K.recon_bwd, emu.recon_bwd):
            dy = torch.full((1, N, B, Tp), 9.0, device=DEV, dtype=torch.bfloat16)
            dg, db, dbi = (torch.empty(N, device=DEV) for _ in range(3))
            fn(y, stats, gamma, beta, None, None, None, inv, ext, dy, dg, db, dbi, T, G, kind)
            outs.append((dy.float(), dg, db, dbi))
        for a, b, nm in zip(outs[0], outs[1], ("dy", "dgamma", "dbeta", "dbias")):
            close(a, b, 5e-3, nm)


@pytest.mark.parametrize("O", [8, 64])
@pytest.mark.parametrize("B,T", [(3, 21), (600, 1)])          # (600, 1): large-batch static fields, gradient staged in chunks
def test_head_fwd_bwd(O, B, T):
    C = 24
    Tp = tp_of(T)
    h = cr(C, B, T, seed=1)
    w = rnd(O, C * T, seed=2, scale=0.1)
    sigma = torch.tensor([1.7], device=DEV)
    bias = rnd(O, seed=3)
    o1, o2 = torch.empty(B, O, device=DEV), torch.empty(B, O, device=DEV)
    K.head_fwd(h, w, sigma, bias, o1, T)
    emu.head_fwd(h, w, sigma, bias, o2, T)
    close(o1, o2, 1e-5, "head")
    dout = rnd(B, O, seed=4)
    res = []
    for fn in (K.head_bwd, emu.head_bwd):
        dwn, dbias = torch.empty(O, C * T, device=DEV), torch.empty(O, device=DEV)
        dh = cr(C, B, T, seed=5)
        fn(h, w, sigma, dout, dwn, dbias, dh, 1, T)
        res.append((dwn, dbias, dh[:, :, :T]))
    for a, b, nm in zip(res[0], res[1], ("dwn", "dbias", "dh")):
        close(a, b, 1e-5, nm)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_latent_fwd_bwd(dtype):
    D, B, T = 8, 3, 21
    Tp = tp_of(T)
    z = rnd(B, D, seed=1)
    w = rnd(D * T, D, seed=2, scale=0.3)
    sigma = torch.tensor([0.8], device=DEV)
    bias = rnd(D * T, seed=3)
    o1 = torch.full((5, D, B, Tp), 9.0, device=DEV, dtype=dtype)
    o2 = torch.empty_like(o1)
    K.latent_fwd(z, w, sigma, bias, o1, T)
    emu.latent_fwd(z, w, sigma, bias, o2, T)
    close(o1.float(), o2.float(), 1e-5 if dtype == torch.float32 else 4e-3, "latent")
    dact = cr(D, B, T, seed=4)
    res = []
    for fn in (K.latent_bwd, emu.latent_bwd):
        dwn, dbias, dz = torch.empty(D * T, D, device=DEV), torch.empty(D * T, device=DEV), torch.empty(B, D, device=DEV)
        fn(z, w, sigma, dact, dwn, dbias, dz, T)
        res.append((dwn, dbias, dz))
    for a, b, nm in zip(res[0], res[1], ("dwn", "dbias", "dz")):
        close(a, b, 1e-5, nm)


def test_reparam_main():
    B, L = 5, 32
    last = rnd(B, 2 * L, seed=1) * 2
    last[0, L] = 40.0      # exercises the clamp
    last[1, L + 1] = -40.0
    last[2, L + 2] = 8.0   # std clamp (exp(4) > 10)
    eps = rnd(B, L, seed=2)
    z1, z2 = torch.empty(B, L, device=DEV), torch.empty(B, L, device=DEV)
    k1, k2 = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    K.reparam_main_fwd(last, eps, z1, k1)
    emu.reparam_main_fwd(last, eps, z2, k2)
    close(z1, z2, 1e-6, "z")
    close(k1, k2, 1e-6, "kl")
    dz, dkl = rnd(B, L, seed=3), torch.tensor([1e-4], device=DEV)
    d1, d2 = torch.empty_like(last), torch.empty_like(last)
    K.reparam_main_bwd(last, eps, dz, dkl, d1)
    emu.reparam_main_bwd(last, eps, dz, dkl, d2)
    close(d1, d2, 1e-5, "dlast")


@pytest.mark.parametrize("std_scale", [1.0, 1e-10])
def test_kl2_reparam(std_scale):
    C, B, T = 16, 3, 21
    Tp = tp_of(T)
    cz, cxz = cr(2 * C, B, T, seed=1) * 1.5, cr(2 * C, B, T, seed=2) * 1.5
    cz[C, 0, 0] = 35.0
    cxz[C + 1, 1, 1] = -35.0
    cz[C + 2, 0, 2] = 20.0
    eps = rnd(B, C, T, seed=3)
    h = cr(C, B, T, seed=4)
    zs1 = torch.full((3, C, B, Tp), 9.0, device=DEV, dtype=torch.bfloat16)
    zs2 = torch.empty_like(zs1)
    f1, f2 = torch.empty(C, B, Tp, device=DEV), torch.empty(C, B, Tp, device=DEV)
    k1, k2 = (torch.empty(1, device=DEV, dtype=torch.float64) for _ in range(2))
    K.kl2_reparam_fwd(cz, cxz, eps, h, std_scale, zs1, f1, k1, T)
    emu.kl2_reparam_fwd(cz, cxz, eps, h, std_scale, zs2, f2, k2, T)
    close(f1, f2, 1e-6, "zs")
    close(zs1.float(), zs2.float(), 4e-3, "zs op")
    close(k1, k2, 1e-5, "kl sum")
    dzs = cr(C, B, T, seed=5)
    dkl = torch.tensor([1e-4], device=DEV)
    res = []
    for fn in (K.kl2_reparam_bwd, emu.kl2_reparam_bwd):
        a, b = torch.full((2 * C, B, Tp), 9.0, device=DEV), torch.full((2 * C, B, Tp), 9.0, device=DEV)
        fn(cz, cxz, eps, std_scale, dzs, dkl, 0.5 / B, a, b, T)
        res.append((a, b))
    close(res[0][0], res[1][0], 1e-5, "dcz")
    close(res[0][1], res[1][1], 1e-5, "dcxz")


def test_philox_normal_statistics_and_batch_split_invariance():
    B, per = 8, 4099
    a = torch.empty(B, per, device=DEV)
    K.philox_normal(a, 1234, 0, 0)
    assert abs(float(a.mean())) < 0.02 and abs(float(a.std()) - 1.0) < 0.02
    assert abs(float((a[0] * a[1]).mean())) < 0.05
    b = torch.empty(B // 2, per, device=DEV)
    K.philox_normal(b, 1234, 0, 4)          # second half of the batch drawn on "another rank"
    assert torch.equal(a[4:], b)
    c = torch.empty(B, per, device=DEV)
    K.philox_normal(c, 1234, 1, 0)          # another draw
    assert not torch.equal(a, c)
    # tails look Gaussian
    frac = float((a.abs() > 1.96).float().mean())
    assert 0.04 < frac < 0.06


def test_adamw_matches_torch():
    n = 10007
    p0, g = rnd(n, seed=1), rnd(n, seed=2)
    p = p0.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3)
    gn = torch.zeros(1, device=DEV, dtype=torch.float64)
    for step in range(1, 4):
        ref.grad = g.clone()
        opt.step()
        K.adamw_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 1.0, gn)
    close(p, ref.detach(), 1e-6, "adamw")
    assert abs(float(gn) - 3 * float((g.double() ** 2).sum())) / float(gn) < 1e-9


def test_opt_step_multi_tensor_matches_model():
    """sg_opt_step (spectral-norm gradient + AdamW + grad norm for all tensors in two launches) against the
    per-tensor torch model, over every layout the fused kernel distinguishes."""
    specs = [("conv", 40, 24, 3), ("convT", 40, 24, 3), ("linear", 16, 64, 1), ("conv", 24, 20, 1), ("conv", 300, 40, 5),
             ("vec", 10007, 0, 0), ("vec", 4096, 0, 0), ("linear", 8, 12000, 1)]

    def make():
        items = []
        for si, (kind, a, b, k) in enumerate(specs):
            if kind == "vec":
                p = rnd(a, seed=si)
                items.append(dict(p=p, g=rnd(a, seed=100 + si), m=rnd(a, seed=200 + si) * 0.1,
                                  v=rnd(a, seed=300 + si).abs() * 0.01))
                continue
            Cout, Cin = a, b
            Cin_p = Cin if kind == "linear" else (Cin + 7) // 8 * 8
            shape = (Cout, Cin) if kind == "linear" else ((Cin, Cout, k) if kind == "convT" else (Cout, Cin, k))
            p = rnd(*shape, seed=si)
            g = rnd(k, Cout, Cin_p, seed=100 + si)
            items.append(dict(p=p, g=g, m=torch.zeros_like(p), v=torch.zeros_like(p),
                              u=torch.nn.functional.normalize(rnd(Cout, seed=400 + si), dim=0),
                              vv=torch.nn.functional.normalize(rnd(Cin * k, seed=500 + si), dim=0),
                              sigma=torch.tensor([1.3 + 0.1 * si], device=DEV), Cout=Cout, Cin=Cin, Cin_p=Cin_p, k=k,
                              flip=int(kind == "convT")))
        return items
    a_items, b_items = make(), make()
    pa, pb = K.OptPlan(a_items, DEV), emu.OptPlan(b_items, DEV)
    ga, gb = (torch.zeros(1, device=DEV, dtype=torch.float64) for _ in range(2))
    for step in (1, 2, 3):
        K.opt_step(pa, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 0.5, ga)
        emu.opt_step(pb, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 0.5, gb)
    for ia, ib, sp in zip(a_items, b_items, specs):
        close(ia["p"], ib["p"], 2e-6, "p %s" % (sp,))
        close(ia["m"], ib["m"], 2e-5, "m %s" % (sp,))
        close(ia["v"], ib["v"], 2e-5, "v %s" % (sp,))
    assert abs(float(ga) - float(gb)) / float(gb) < 1e-5


@pytest.mark.parametrize("where", ["weight", "vector"])
def test_opt_step_dynamic_loss_scaler_skips_overflowed_steps(where):
    """sg_scaler_state: clean steps divide the loss scale out and count; a step with a non-finite gradient (in a
    spectral-norm weight gradient or in a plain vector) leaves p / m / v / step untouched, halves the scale and reports
    an infinite gradient norm; `growth_interval` clean steps double the scale.  Kernel against the torch model."""
    def make():
        p1, p2 = rnd(24, 20, 3, seed=1), rnd(4099, seed=2)
        return [dict(p=p1, g=rnd(3, 24, 24, seed=3) * 64.0, m=torch.zeros_like(p1), v=torch.zeros_like(p1),
                     u=torch.nn.functional.normalize(rnd(24, seed=4), dim=0),
                     vv=torch.nn.functional.normalize(rnd(60, seed=5), dim=0), sigma=torch.tensor([1.7], device=DEV),
                     Cout=24, Cin=20, Cin_p=24, k=3, flip=0),
                dict(p=p2, g=rnd(4099, seed=6) * 64.0, m=torch.zeros_like(p2), v=torch.zeros_like(p2))]
    a_items, b_items = make(), make()
    pa, pb = K.OptPlan(a_items, DEV), emu.OptPlan(b_items, DEV)
    sa = K.make_scaler_state(DEV, 64.0, growth_interval=2)
    sb = K.make_scaler_state(DEV, 64.0, growth_interval=2)
    ga, gb = (torch.zeros(1, device=DEV, dtype=torch.float64) for _ in range(2))
    bad_item = 0 if where == "weight" else 1
    for it in range(5):
        if it == 1:                                          # overflow in this step's gradients
            for items in (a_items, b_items):
                items[bad_item]["g"].view(-1)[17] = float("inf") if where == "weight" else float("nan")
        if it == 2:
            for items in (a_items, b_items):
                items[bad_item]["g"].view(-1)[17] = 1.0
        before = [x["p"].clone() for x in a_items]
        ga.zero_(); gb.zero_()
        K.opt_step(pa, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0, 1.0, ga, sa)
        emu.opt_step(pb, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0, 1.0, gb, sb)
        st = K.read_scaler_state(sa)
        assert st == K.read_scaler_state(sb), (it, st, K.read_scaler_state(sb))
        if it == 1:
            assert st["last_skipped"] == 1 and st["skipped"] == 1 and st["step"] == 1 and st["scale"] == 32.0
            assert float(ga) == float("inf")
            for x, b in zip(a_items, before):
                assert torch.equal(x["p"], b)
        else:
            assert abs(float(ga) - float(gb)) / float(gb) < 1e-5
    st = K.read_scaler_state(sa)
    assert st["step"] == 4 and st["skipped"] == 1 and st["scale"] == 64.0     # 64 -> 32 (skipped step) -> 64 (two clean steps in a row)
    for ia, ib in zip(a_items, b_items):
        close(ia["p"], ib["p"], 2e-6, "p")
        close(ia["m"], ib["m"], 2e-5, "m")
        close(ia["v"], ib["v"], 2e-5, "v")


@pytest.mark.parametrize("loss", ["MSE", "MAE", "smoothL1", "Huber"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("with_xhat", [False, True])
@pytest.mark.parametrize("T", [200, 8])                    # 8: short-row kernels
def test_recon_head_with_packed_operand_target(loss, dtype, with_xhat, T):
    """The reconstruction head reading its target from the packed 16-bit operand of x ([N, B, Tp], the layout of y;
    engine.loss_target() == "operand") instead of the fp32 [B, N, T] tensor: forward sums, optional x_hat and the one-pass
    backward against the torch model fed the same rounded target."""
    N, B, G = (264, 5, 8) if T > 8 else (72, 70, 8)
    Tp = engine.tp_of(T, "bf16")
    kind = K.LOSS_KINDS[loss]
    y = (rnd(N, B, Tp, seed=1) * 2.0).to(dtype)
    gamma, beta = rnd(N, seed=2) * 0.5 + 1.0, rnd(N, seed=3) * 0.2
    stats = torch.empty(B, G, 2, device=DEV)
    K.gn_stats(y.float(), stats, T, G)
    x = (rnd(B, N, T, seed=4) * 0.5).clamp(-0.7, 0.7)
    op = torch.empty(1, N, B, Tp, device=DEV, dtype=dtype)
    K.pack_input(x, op, T)
    xh1 = torch.empty(B, N, T, device=DEV) if with_xhat else None
    xh2 = torch.empty(B, N, T, device=DEV)
    s1, s2 = (torch.empty(2, device=DEV, dtype=torch.float64) for _ in range(2))
    rowsums = torch.full((N * B, 4), 9.0, device=DEV)
    K.recon_fwd(y, stats, gamma, beta, op[0], xh1, s1, T, G, kind, rowsums)
    emu.recon_fwd(y, stats, gamma, beta, op[0], xh2, s2, T, G, kind)
    close(s1, s2, 1e-5, "loss sums")
    if with_xhat:
        close(xh1, xh2, 3e-6, "x_hat")
    # against the fp32 target the loss moves by the rounding of x only
    s3 = torch.empty(2, device=DEV, dtype=torch.float64)
    emu.recon_fwd(y, stats, gamma, beta, x, xh2, s3, T, G, kind)
    close(s1, s3, 2e-4 if dtype == torch.float16 else 2e-3, "loss sums vs fp32 target")
    g_loss, g_mse = torch.tensor([1.0e6], device=DEV), torch.tensor([0.5], device=DEV)
    inv = 1.0 / (B * N * T)
    outs = []
    for fn in (K.recon_bwd, emu.recon_bwd):
        dy = torch.full((1, N, B, Tp), 9.0, device=DEV, dtype=dtype)
        dg, db, dbi = (torch.empty(N, device=DEV) for _ in range(3))
        fn(y, stats, gamma, beta, op[0], g_loss, g_mse, inv, None, dy, dg, db, dbi, T, G, kind, rowsums)
        outs.append((dy.float(), dg, db, dbi))
    tol16 = 6e-4 if dtype == torch.float16 else 5e-3
    for a, b, nm in zip(outs[0], outs[1], ("dy", "dgamma", "dbeta", "dbias")):
        close(a, b, tol16 if nm in ("dy", "dbias") else 5e-5, nm)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_sn_prepare_batched_matches_per_layer_model(training, dtype):
    """sg_sn_prepare (all layers in five launches) against the per-layer torch model."""
    specs = [("conv", 40, 24, 3, True), ("convT", 40, 24, 3, True), ("linear", 16, 4100, 1, True), ("conv", 72, 64, 1, True),
             ("conv", 24, 20, 5, False), ("conv", 300, 9000, 1, True), ("linear", 2000, 8, 1, True)]

    def make():
        layers = []
        for si_, (kind, Cout, Cin, k, sn) in enumerate(specs):
            if kind == "linear":
                w, so, si, flip, Cin_p = rnd(Cout, Cin, seed=si_), Cin, 1, 0, Cin
            elif kind == "convT":
                w, so, si, flip, Cin_p = rnd(Cin, Cout, k, seed=si_), k, Cout * k, 1, (Cin + 7) // 8 * 8
            else:
                w, so, si, flip, Cin_p = rnd(Cout, Cin, k, seed=si_), Cin * k, k, 0, (Cin + 7) // 8 * 8
            L = dict(w=w, sigma=torch.zeros(1, device=DEV), H=Cout, Cin=Cin, k=k, Cin_p=Cin_p, so=so, si=si, flip=flip,
                     wg=None if kind == "linear" else torch.full((k, Cout, Cin_p), 3.0, device=DEV, dtype=dtype))
            if sn:
                L["u"] = torch.nn.functional.normalize(rnd(Cout, seed=50 + si_), dim=0)
                L["v"] = torch.nn.functional.normalize(rnd(Cin * k, seed=80 + si_), dim=0)
            layers.append(L)
        return layers
    la, lb = make(), make()
    pa, pb = K.SnPlan(la, DEV, dtype), emu.SnPlan(lb, DEV, dtype)
    for _ in range(2):
        K.sn_prepare(pa, training)
        emu.sn_prepare(pb, training)
    for A, Bm, sp in zip(la, lb, specs):
        close(A["sigma"], Bm["sigma"], 1e-5, "sigma %s" % (sp,))
        if sp[4]:
            close(A["u"], Bm["u"], 1e-5, "u %s" % (sp,))
            close(A["v"], Bm["v"], 1e-5, "v %s" % (sp,))
        if A["wg"] is not None:
            close(A["wg"].float(), Bm["wg"].float(), 1e-6 if dtype == torch.float32 else 4e-3, "wg %s" % (sp,))


# ---- static fields (T = 1): compact [C][B] forms of the two N-channel layers (csrc/static_ops.cu) ------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,N", [(8, 37), (72, 100), (512, 1000)])
@pytest.mark.parametrize("with_xt", [False, True])
def test_pack_static(dtype, B, N, with_xt):
    x = rnd(B, N, 1, seed=1)
    outs = []
    for fn in (K.pack_static, emu.pack_static):
        xc = torch.full((1, N, B // 8, 8), 7.0, device=DEV, dtype=dtype)
        xt = torch.full((N, B), 7.0, device=DEV) if with_xt else None
        fn(x, xc, xt)
        outs.append((xc, xt))
    assert torch.equal(outs[0][0], outs[1][0])
    if with_xt:
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][1], x[:, :, 0].t())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_rows_compact_expand(dtype):
    C, B = 48, 72
    padded = rnd(C, B, 8, seed=2).to(dtype)
    c1, c2 = (torch.full((C, B // 8, 8), 9.0, device=DEV, dtype=dtype) for _ in range(2))
    K.rows_compact16(padded, c1)
    emu.rows_compact16(padded, c2)
    assert torch.equal(c1, c2) and torch.equal(c1.view(C, B), padded[:, :, 0])
    comp = rnd(C, B // 8, 8, seed=3)
    for acc in (False, True):
        p1 = cr(C, B, 1, seed=4)
        p2 = p1.clone()
        K.rows_expand_f32(comp, p1, acc)
        emu.rows_expand_f32(comp, p2, acc)
        assert torch.equal(p1, p2)
        assert float(p1[:, :, 1:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,N,G", [(8, 64, 8), (24, 1000, 8), (256, 1000, 8), (512, 4096, 8), (2048, 96, 4)])
def test_static_stats(dtype, B, N, G):
    y = (rnd(N, B, seed=1) * 1.5 + 0.3).to(dtype)
    s1, s2 = torch.empty(B, G, 2, device=DEV), torch.empty(B, G, 2, device=DEV)
    K.static_stats(y, s1, G)
    emu.static_stats(y, s2, G)
    close(s1, s2, 2e-6, "stats (mean, rstd)")


@pytest.mark.parametrize("loss", ["MSE", "MAE", "smoothL1", "Huber"])
@pytest.mark.parametrize("dtype,x16", [(torch.float16, True), (torch.float16, False), (torch.bfloat16, False)])
@pytest.mark.parametrize("B,N", [(8, 64), (24, 1000), (256, 1000), (512, 4096)])
def test_static_recon_fwd_bwd(loss, dtype, x16, B, N):
    """The compact reconstruction head against the torch model of the padded one (T = 1): loss sums, dy, dgamma, dbeta,
    dbias.  B = 24: 3 octets per channel (warps straddle channels: per-thread atomics); B = 256 / 512: whole warps."""
    G = 8
    kind = K.LOSS_KINDS[loss]
    y = (rnd(N, B, seed=1) * 2.0).to(dtype)
    gamma, beta = rnd(N, seed=2) * 0.5 + 1.0, rnd(N, seed=3) * 0.2
    x = (rnd(N, B, seed=4) * 0.5).clamp(-0.7, 0.7)
    if x16:
        x = x.to(dtype)
    stats = torch.empty(B, G, 2, device=DEV)
    K.static_stats(y, stats, G)
    g_loss, g_mse = torch.tensor([1.0e6], device=DEV), torch.tensor([0.5], device=DEV)
    inv = 1.0 / (B * N)
    outs = []
    for mod in (K, emu):
        sums = torch.empty(2, device=DEV, dtype=torch.float64)
        ws = mod.static_recon_ws(N, B, G, DEV)
        xh = torch.full((N, B), 9.0, device=DEV)
        mod.static_recon_fwd(y, stats, gamma, beta, x, sums, ws, G, kind, xh)
        dy = torch.full((N, B), 9.0, device=DEV, dtype=dtype)
        dg, db, dbi = (torch.full((N,), 9.0, device=DEV) for _ in range(3))
        mod.static_recon_bwd(y, stats, gamma, beta, x, g_loss, g_mse, inv, ws, dy, dg, db, dbi, G, kind)
        outs.append((sums, dy.float(), dg, db, dbi, xh))
    close(outs[0][0], outs[1][0], 1e-5, "loss sums")
    close(outs[0][5], outs[1][5], 3e-6, "x_hat (transposed)")
    tol16 = 6e-4 if dtype == torch.float16 else 5e-3          # dy is stored in the 16-bit operand format
    for a, b, nm in zip(outs[0][1:5], outs[1][1:5], ("dy", "dgamma", "dbeta", "dbias")):
        close(a, b, tol16 if nm in ("dy", "dbias") else 5e-5, nm)
