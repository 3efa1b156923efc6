"""-m gpu: the tcgen05/TMEM/TMA implicit-GEMM convolution kernels (bf16 operands, fp32 accumulate)
against the plain PyTorch conv of the same bf16 inputs, through the C ABI."""
import pytest
import torch

import kernel_emulator as emu
from conftest import rel_l2
from simulgen_vae_b200 import kernels as K
from simulgen_vae_b200 import engine


def tp_of(T):
    """bf16-mode row pitch: no zero gap between samples (the taps read pre-shifted planes)"""
    return engine.tp_of(T, "bf16")

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16

SHAPES = [
    # Cin, Cout, k, B, T
    (24, 72, 1, 3, 21),          # everything ragged, single tile
    (24, 72, 3, 3, 21),
    (64, 128, 5, 2, 30),
    (256, 384, 3, 4, 200),       # multi-tile M and N, T = 200 -> Tp = 208
    (1280, 1280, 5, 2, 200),     # the decoder residual shape (scaled), deep K loop
    (8200, 136, 1, 2, 100),      # large K with split-K, ragged M
    (200, 4104, 1, 2, 100),      # recon-like: large M
    (300, 2600, 3, 6, 200),      # CTA-pair kernel: 11 pair m-tiles (raster groups of 8 + 3) x 5 n-tiles, ragged everywhere
    (136, 300, 1, 1, 40),        # pair tile with a mostly empty peer half, single n-tile
]


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def make(Cin, Cout, k, B, T):
    Tp = tp_of(T)
    Cin_p = (Cin + 7) // 8 * 8
    wg = rnd(k, Cout, Cin_p, seed=1, scale=1.0 / (Cin * k) ** 0.5)
    wg[:, :, Cin:] = 0
    # operands as the engine stores them: k pre-shifted planes (one more pair than needed for k=1, to
    # exercise the plane offset arithmetic)
    Pa = 3 if k == 1 else k
    act = torch.empty(Pa, Cin, B, Tp, device=DEV, dtype=BF)
    emu.write_planes(act, rnd(Cin, B, T, seed=2), T)
    dy = torch.empty(k, Cout, B, Tp, device=DEV, dtype=BF)
    emu.write_planes(dy, rnd(Cout, B, T, seed=3), T)
    return wg.to(BF), act, dy, rnd(Cout, seed=4), Tp, Cin_p


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_fprop(shape):
    Cin, Cout, k, B, T = shape
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    o1 = torch.full((Cout, B, Tp), 5.0, device=DEV)
    o2 = torch.empty_like(o1)
    K.conv_fprop(wg, act, bias, o1, Cin)
    emu.conv_fprop(wg, act, bias, o2, Cin)
    torch.cuda.synchronize()
    # gap columns t >= T are don't-care (the model convolves across the zero gap, the planes do not)
    assert rel_l2(o1[:, :, :T], o2[:, :, :T]) < 1e-4, rel_l2(o1[:, :, :T], o2[:, :, :T])
    K.conv_fprop(wg, act, None, o1, Cin, accumulate=True)
    emu.conv_fprop(wg, act, None, o2, Cin, accumulate=True)
    assert rel_l2(o1[:, :, :T], o2[:, :, :T]) < 1e-4, rel_l2(o1[:, :, :T], o2[:, :, :T])


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_dgrad(shape):
    Cin, Cout, k, B, T = shape
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    d1 = torch.full((Cin, B, Tp), 5.0, device=DEV)
    d2 = torch.empty_like(d1)
    K.conv_dgrad(wg, dy, d1, Cin)
    emu.conv_dgrad(wg, dy, d2, Cin)
    torch.cuda.synchronize()
    assert rel_l2(d1[:, :, :T], d2[:, :, :T]) < 1e-4, rel_l2(d1[:, :, :T], d2[:, :, :T])
    K.conv_dgrad(wg, dy, d1, Cin, accumulate=True)
    emu.conv_dgrad(wg, dy, d2, Cin, accumulate=True)
    assert rel_l2(d1[:, :, :T], d2[:, :, :T]) < 1e-4, rel_l2(d1[:, :, :T], d2[:, :, :T])


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[0] > 128] + [(2560, 512, 1, 3, 200), (1280, 1280, 5, 1, 8)])
@pytest.mark.parametrize("half", [False, True])
def test_tc_dgrad_16bit_output(shape, half):
    """dgrad with the gradient stored in the 16-bit operand format by the CTA-pair epilogue (activations whose only
    consumer is this conv), plain and accumulating onto a 16-bit buffer, for bf16 and fp16 operands."""
    Cin, Cout, k, B, T = shape
    assert K.conv_out16_ok(Cin)
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    dt = torch.float16 if half else BF
    if half:
        wg, dy = wg.float().half(), dy.float().half()
    ref = torch.empty(Cin, B, Tp, device=DEV)
    emu.conv_dgrad(wg, dy, ref, Cin)
    d1 = torch.full((Cin, B, Tp), 5.0, device=DEV, dtype=dt)
    K.conv_dgrad(wg, dy, d1, Cin)
    torch.cuda.synchronize()
    tol = 6e-4 if half else 4e-3
    assert rel_l2(d1[:, :, :T].float(), ref[:, :, :T]) < tol, rel_l2(d1[:, :, :T].float(), ref[:, :, :T])
    d2 = d1.clone()
    K.conv_dgrad(wg, dy, d1, Cin, accumulate=True)
    emu.conv_dgrad(wg, dy, d2, Cin, accumulate=True)
    torch.cuda.synchronize()
    assert rel_l2(d1[:, :, :T].float(), d2[:, :, :T].float()) < tol, rel_l2(d1[:, :, :T].float(), d2[:, :, :T].float())
    assert rel_l2(d1[:, :, :T].float(), 2 * ref[:, :, :T]) < tol * 1.5


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_wgrad(shape):
    Cin, Cout, k, B, T = shape
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    w1 = torch.full((k, Cout, Cin_p), 5.0, device=DEV)
    w2 = torch.empty_like(w1)
    K.conv_wgrad(dy, act, w1, Cin)
    emu.conv_wgrad(dy, act, w2, Cin)
    torch.cuda.synchronize()
    assert rel_l2(w1[:, :, :Cin], w2[:, :, :Cin]) < 1e-4, rel_l2(w1[:, :, :Cin], w2[:, :, :Cin])
    if Cin_p > Cin:
        assert float(w1[:, :, Cin:].abs().max()) == 0.0


@pytest.mark.parametrize("shape", SHAPES + [(64, 320, 1, 5, 200), (96, 1000, 3, 3, 8)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_tc_fprop_gn_fused_statistics(shape, out_bf16):
    """conv + GroupNorm statistics in one call: (mean, rstd) from the GEMM epilogue (CTA-pair kernel) or from the
    separate pass (single-CTA / split-K) must equal the statistics of the fp32 conv output; optional bf16 output."""
    Cin, Cout, k, B, T = shape
    G = 8
    if Cout % G or (out_bf16 and Cout <= 128):
        pytest.skip("not a GroupNorm / bf16-output shape")
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    o2 = torch.empty(Cout, B, Tp, device=DEV)
    emu.conv_fprop(wg, act, bias, o2, Cin)
    o2[:, :, T:] = 0
    s2 = torch.empty(B, G, 2, device=DEV)
    emu.gn_stats(o2, s2, T, G)
    o1 = torch.full((Cout, B, Tp), 5.0, device=DEV, dtype=BF if out_bf16 else torch.float32)
    s1 = torch.full((B, G, 2), 7.0, device=DEV)
    try:
        K.conv_fprop_gn(wg, act, bias, o1, Cin, s1, T, G)
    except RuntimeError as e:
        if out_bf16 and "fused-statistics" in str(e):
            pytest.skip("split-K shape: bf16 output not available")
        raise
    torch.cuda.synchronize()
    assert rel_l2(o1[:, :, :T].float(), o2[:, :, :T]) < (4e-3 if out_bf16 else 1e-4)
    assert rel_l2(s1[:, :, 0], s2[:, :, 0]) < 1e-3 and (s1[:, :, 0] - s2[:, :, 0]).abs().max() < 1e-4
    assert rel_l2(s1[:, :, 1], s2[:, :, 1]) < 1e-4


@pytest.mark.parametrize("shape", [(256, 384, 3, 4, 200), (1280, 1280, 5, 2, 200), (24, 72, 3, 3, 21)])
def test_tc_fp16_operands(shape):
    """'fp16' precision mode: the same tcgen05 kernels with IEEE fp16 operand formats (kind::f16, fp32 accumulation)."""
    Cin, Cout, k, B, T = shape
    wg, act, dy, bias, Tp, Cin_p = make(*shape)
    wg, act, dy = wg.float().half(), act.float().half(), dy.float().half()
    o1, o2 = torch.full((Cout, B, Tp), 5.0, device=DEV), torch.empty(Cout, B, Tp, device=DEV)
    K.conv_fprop(wg, act, bias, o1, Cin)
    emu.conv_fprop(wg, act, bias, o2, Cin)
    assert rel_l2(o1[:, :, :T], o2[:, :, :T]) < 1e-4
    d1, d2 = torch.empty(Cin, B, Tp, device=DEV), torch.empty(Cin, B, Tp, device=DEV)
    K.conv_dgrad(wg, dy, d1, Cin)
    emu.conv_dgrad(wg, dy, d2, Cin)
    assert rel_l2(d1[:, :, :T], d2[:, :, :T]) < 1e-4
    w1, w2 = torch.empty(k, Cout, Cin_p, device=DEV), torch.empty(k, Cout, Cin_p, device=DEV)
    K.conv_wgrad(dy, act, w1, Cin)
    emu.conv_wgrad(dy, act, w2, Cin)
    assert rel_l2(w1[:, :, :Cin], w2[:, :, :Cin]) < 1e-4
    # back to bf16 in the same process: the format is selected per call
    wgb, actb, _, _, _, _ = make(*shape)
    K.conv_fprop(wgb, actb, bias, o1, Cin)
    emu.conv_fprop(wgb, actb, bias, o2, Cin)
    assert rel_l2(o1[:, :, :T], o2[:, :, :T]) < 1e-4
