"""Standalone calls of the overlay's block modules (`ResidualBlock(dim, small)(x)`, `EncoderBlock`, `UpsampleBlock`,
...): the reference's blocks are ordinary nn.Modules that can be called on a [B, C, T] tensor
(/root/reference/modules/common.py:101-102,124-125,161-162, encoder.py:45-46,91-94, decoder.py:32-33,79-82).  Inside
VAE.forward the engine never calls them one by one (the encoder / decoder are each ONE autograd Function); this module
gives the same classes a working `forward` of their own, through the same kernels: x is packed into the engine's CR
layout, the block's layers run as fused conv + GroupNorm + activation steps, the result is unpacked to [B, C', T].
Gradients flow to x and to the block's parameters (spectral-norm hooks included)."""
from __future__ import annotations

import torch

from . import engine
from . import kernels as K


def _first_conv(block):
    kind = block._sg_kind
    if kind == "chain":
        return _first_conv(block.module_list[0])
    seq = block._seq if hasattr(block, "_seq") else block.seq
    return seq[0]


def _run(ctx, block, a, out_planes):
    """One block on the activation `a`; the result carries `out_planes` operand planes and an fp32 copy."""
    kind = block._sg_kind
    if kind == "chain":
        mods = list(block.module_list)
        for i, sub in enumerate(mods):
            nxt = engine._k(_first_conv(mods[i + 1])) if i + 1 < len(mods) else out_planes
            a = _run(ctx, sub, a, nxt)
        return a
    seq = block._seq if hasattr(block, "_seq") else block.seq
    if kind == "plain":                                     # ConvBlock: (conv, GN, GELU) x 1 or 2
        return engine.cgg_seq(ctx, seq, a, want_f32=True, out_planes=out_planes)
    if kind == "residual":                                  # x + 0.1 * seq(x)
        return engine.cgg_seq(ctx, seq, a, res=a, res_scale=0.1, want_f32=True, out_planes=out_planes)
    if kind == "upsample":                                  # ConvTranspose1d - GELU, no norm
        return engine.conv_block(ctx, seq[0], None, a, K.ACT_GELU, want_f32=True, out_planes=out_planes, transposed=True)
    raise RuntimeError("simulgen_b200: unknown block kind %r" % (kind,))


class BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, block, x, *params):
        engine._check_input(x, "block input")
        fctx.set_materialize_grads(False)
        with engine.device_guard(x.device):
            x = x.contiguous().float()
            B, C, T = x.shape
            record = engine._take_grad_mode() and any(fctx.needs_input_grad)
            ctx = engine.Ctx(B, T, x.device, record)
            k = engine._k(_first_conv(block))
            xf = ctx.f32(1, C, B, ctx.Tp)
            K.pack_input(x, xf, T)                          # [B, C, T] -> CR layout, fp32
            a = engine.Act(C, data=ctx.op(k, C, B, ctx.Tp), f32=xf[0], needs_grad=fctx.needs_input_grad[1], name="x")
            K.gn_act_fwd(xf[0], None, None, None, None, 1.0, K.ACT_NONE, False, a.data, None, T, 0)   # operand planes
            out = _run(ctx, block, a, 1)
            src = out.as_f32()
            y = torch.empty(B, out.C, T, dtype=torch.float32, device=x.device)
            K.unpack_f32(src, y, T)
        fctx.ectx, fctx.a, fctx.out, fctx.params = ctx, a, out, params
        return y

    @staticmethod
    def backward(fctx, gy):
        ctx, a, out = fctx.ectx, fctx.a, fctx.out
        if gy is None:
            return (None, None) + tuple(None for _ in fctx.params)
        with engine.device_guard(ctx.dev):
            gy = engine._contig_f32(gy)
            g = ctx.f32(1, out.C, ctx.B, ctx.Tp)
            K.pack_input(gy, g, ctx.T)
            out.grad = g[0]
            ctx.run_backward()
            gx = None
            if a.needs_grad and a.grad is not None:
                gx = torch.empty(ctx.B, a.C, ctx.T, dtype=torch.float32, device=ctx.dev)
                K.unpack_f32(a.grad, gx, ctx.T)
        pg = tuple(ctx.pgrads.get(id(p)) for p in fctx.params)
        fctx.ectx = None
        return (None, gx) + pg


def run_block(block, x):
    engine.note_grad_mode()
    return BlockFn.apply(block, x, *block.parameters())
