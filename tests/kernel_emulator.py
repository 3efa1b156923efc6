"""TEST INFRASTRUCTURE ONLY - plain PyTorch fp32 models of every C-ABI kernel.

Two uses:
  * `-m gpu` kernel tests compare each CUDA kernel with the function of the same name here
    (the "plain PyTorch reference of the same op"); the backward models use torch.autograd on the
    forward formula, so they check the hand-derived gradients independently.
  * CPU tests install these functions over `simulgen_vae_b200.kernels` (see `install()`) to check the
    host-side wiring (tape, layouts, spectral-norm bookkeeping, autograd boundary) against the oracle
    without a GPU.  The product never does this: `simulgen_vae_b200.kernels` raises without CUDA.
"""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_TANH = 0, 1, 2


def _act(kind, x):
    if kind == ACT_GELU:
        return F.gelu(x)
    if kind == ACT_TANH:
        return torch.tanh(x)
    return x


# ---- operand planes -----------------------------------------------------------------------------
def write_planes(dst, val, T):
    """dst [P, C, B, Tp] <- val [C, B, T]: dst[pl][c][b][t] = val[c][b][t + pl - P//2], zero outside [0, T)."""
    P = dst.shape[0]
    dst.zero_()
    for pl in range(P):
        s = pl - P // 2
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            dst[pl, :, :, lo:hi] = val[:, :, lo + s:hi + s].to(dst.dtype)


def center(t):
    return t[t.shape[0] // 2]


# ---- layout -------------------------------------------------------------------------------------
def pack_input(x, out, T):
    write_planes(out, x.float().permute(1, 0, 2), T)


def unpack_f32(inp, out, T):
    out.copy_(inp[:, :, :T].permute(1, 0, 2))


def axpy(dst, src, alpha, accumulate):
    if accumulate:
        dst.add_(src, alpha=alpha)
    else:
        dst.copy_(src * alpha)


def scale_f64_to_f32(inp, out, scale):
    out.copy_((inp * scale).to(torch.float32))


# ---- spectral norm ------------------------------------------------------------------------------
def _wmat(w, H, Cin, k, so, si):
    if w.dim() == 2 or (so == Cin * k and si == k):
        return w.reshape(H, -1)
    return w.permute(1, 0, 2).reshape(H, -1)          # ConvTranspose1d, dim=1


def sn_power_iter(w, u, v, sigma, H, Cin, k, so, si, training):
    wm = _wmat(w.detach(), H, Cin, k, so, si)
    if training:
        vn = F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
        un = F.normalize(torch.mv(wm, vn), dim=0, eps=1e-12)
        u.copy_(un)
        v.copy_(vn)
    sigma.copy_(torch.dot(u, torch.mv(wm, v)).reshape(1))


def sn_pack_weight(w, sigma, wg, Cout, Cin, Cin_p, k, so, si, flip):
    wn = w.detach() / sigma
    if flip:                                          # [Cin, Cout, k] -> [k(flipped), Cout, Cin]
        g = wn.permute(2, 1, 0).flip(0)
    else:                                             # [Cout, Cin, k] -> [k, Cout, Cin]
        g = wn.permute(2, 0, 1)
    wg.zero_()
    wg[:, :, :Cin] = g.to(wg.dtype)


def sn_weight_grad(dwg, w, u, v, sigma, grad, Cout, Cin, Cin_p, k, so, si, flip):
    w = w.detach()
    if w.dim() == 2:
        G = dwg[0, :, :Cin]
        uv = torch.outer(u, v)
    elif flip:
        G = dwg.flip(0)[:, :, :Cin].permute(2, 1, 0)
        uv = torch.outer(u, v).reshape(Cout, Cin, k).permute(1, 0, 2)
    else:
        G = dwg[:, :, :Cin].permute(1, 2, 0)
        uv = torch.outer(u, v).reshape(Cout, Cin, k)
    dot = (G.double() * w.double()).sum().float()
    grad.copy_((G - dot / sigma * uv) / sigma)


# ---- convolutions -------------------------------------------------------------------------------
# The models use ONLY the centre plane and do the tap shifts themselves (F.conv1d), so they also check
# that the pre-shifted planes the CUDA path reads are consistent with a real convolution.
def _valid_T(t):
    """number of valid columns of a CR operand: the trailing all-zero columns of every row are the gap"""
    return t.shape[-1]


def conv_fprop(wg, act, bias, out, Cin, accumulate=False, T=None):
    """Per-sample convolution of the centre plane (independent of the zero gap between samples); the gap
    columns of `out` receive the same values a flattened convolution over zero-padded rows would give."""
    k, Cout, Cin_p = wg.shape
    a = center(act).float().permute(1, 0, 2)                                  # [B, Cin, Tp]
    w = wg[:, :, :Cin].float().permute(1, 2, 0).contiguous()                  # [Cout, Cin, k]
    y = F.conv1d(a, w, bias.detach() if bias is not None else None, padding=k // 2).permute(1, 0, 2)
    y = y.reshape(out.shape)
    if accumulate:
        out.add_(y)
    else:
        out.copy_(y)


def conv_fprop16(wg, act, bias, out, Cin):
    tmp = torch.empty(out.shape, dtype=torch.float32, device=out.device)
    conv_fprop(wg, act, bias, tmp, Cin)
    out.copy_(tmp.to(out.dtype))


def conv_fprop_gn(wg, act, bias, out, Cin, stats, T, G):
    tmp = torch.empty(out.shape, dtype=torch.float32, device=out.device)
    conv_fprop(wg, act, bias, tmp, Cin)
    gn_stats(tmp, stats, T, G)
    out.copy_(tmp.to(out.dtype))


def conv_dgrad(wg, dy, dx, Cin, accumulate=False):
    k, Cout, Cin_p = wg.shape
    g = center(dy).float().permute(1, 0, 2)                                   # [B, Cout, Tp]
    w = wg[:, :, :Cin].float().flip(0).permute(2, 1, 0).contiguous()          # [ci][co][j'] = wg[k-1-j'][co][ci]
    y = F.conv1d(g, w, None, padding=k // 2).permute(1, 0, 2).reshape(dx.shape)
    if accumulate:
        dx.copy_((dx.float() + y).to(dx.dtype))            # 16-bit dx: add in fp32, round once more
    else:
        dx.copy_(y)


def conv_out16_ok(M):
    return M > 128


def set_sm_limit(sms):
    pass


def conv_wgrad(dy, act, dwg, Cin):
    k, Cout, Cin_p = dwg.shape
    g = center(dy).float()                                                    # [Cout, B, Tp]
    a = center(act).float()                                                   # [Cin, B, Tp]
    Tp = a.shape[2]
    pad = k // 2
    ap = F.pad(a, (pad, pad))                                                 # per-sample zero padding
    dwg.zero_()
    for j in range(k):
        dwg[j, :, :Cin] = torch.einsum("obt,ibt->oi", g, ap[:, :, j:j + Tp])


# ---- GroupNorm + activation ---------------------------------------------------------------------
def gn_stats(y, stats, T, G):
    C, B, Tp = y.shape
    v = y[:, :, :T].double().reshape(G, C // G, B, T)
    mean = v.mean(dim=(1, 3))
    var = (v * v).mean(dim=(1, 3)) - mean * mean
    stats[:, :, 0] = mean.t().float()
    stats[:, :, 1] = (1.0 / torch.sqrt(var.clamp_min(0) + 1e-5)).t().float()


def _gn_forward(y, gamma, beta, res, res_scale, act, post_gelu, T, G, use_gn):
    """y [C,B,Tp] fp32 (differentiable); returns pre/out on the valid region [C,B,T]."""
    C, B, Tp = y.shape
    yv = y[:, :, :T].float()
    if use_gn:
        g = yv.reshape(G, C // G, B, T)
        mean = g.mean(dim=(1, 3), keepdim=True)
        var = g.var(dim=(1, 3), unbiased=False, keepdim=True)
        xh = ((g - mean) / torch.sqrt(var + 1e-5)).reshape(C, B, T)
        yh = xh * gamma[:, None, None] + beta[:, None, None]
    else:
        yh = yv
    pre = res_scale * _act(act, yh)
    if res is not None:
        pre = pre + res[:, :, :T].float()
    return F.gelu(pre) if post_gelu else pre


def gn_act_fwd(y, stats, gamma, beta, res, res_scale, act, post_gelu, out_op, out_f32, T, G):
    with torch.no_grad():
        o = _gn_forward(y, gamma, beta, res, res_scale, act, post_gelu, T, G, stats is not None)
    if out_op is not None:
        write_planes(out_op, o, T)
    if out_f32 is not None:
        out_f32.zero_()
        out_f32[:, :, :T] = o


def gn_act_bwd(y, stats, gamma, beta, res, res_scale, act, post_gelu, dout, dy, dgamma, dbeta, dbias, dres,
               dres_accumulate, T, G, ws=None):
    use_gn = stats is not None
    with torch.enable_grad():
        yl = y.detach().float().clone().requires_grad_(True)
        gl = gamma.detach().clone().requires_grad_(True) if use_gn else None
        bl = beta.detach().clone().requires_grad_(True) if use_gn else None
        rl = res.detach().float().clone().requires_grad_(True) if res is not None else None
        o = _gn_forward(yl, gl, bl, rl, res_scale, act, post_gelu, T, G, use_gn)
        leaves = [t for t in (yl, gl, bl, rl) if t is not None]
        grads = torch.autograd.grad(o, leaves, dout[:, :, :T].float(), allow_unused=True)
    gmap = dict(zip([id(t) for t in leaves], grads))
    gy = gmap[id(yl)]
    gy = gy.clone()
    gy[:, :, T:] = 0
    write_planes(dy, gy[:, :, :T], T)
    if dbias is not None:
        dbias.copy_(gy.sum(dim=(1, 2)))
    if use_gn:
        dgamma.copy_(gmap[id(gl)])
        dbeta.copy_(gmap[id(bl)])
    if dres is not None and rl is not None:
        gr = gmap[id(rl)]
        if dres_accumulate:
            dres[:, :, :T] += gr[:, :, :T]
        else:
            dres.zero_()
            dres[:, :, :T] = gr[:, :, :T]
    if use_gn:
        dout.fill_(float("nan"))       # the CUDA path overwrites dout (with dz): nothing may read it afterwards


def _loss_terms(kind, d):
    if kind == 1:
        return d.abs()
    if kind in (2, 3):
        return torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5)
    return d * d


def _recon_xhat(y, gamma, beta, T, G):
    y = y.float() if y.dtype != torch.float32 else y
    N, B, Tp = y.shape
    o = _gn_forward(y, gamma, beta, None, 1.0, ACT_TANH, False, T, G, True)      # [N,B,T]
    return o.permute(1, 0, 2)                                                     # [B,N,T]


def _target(x, T):
    """fp32 [B, N, T] as is; the packed 16-bit operand plane [N, B, Tp] -> fp32 [B, N, T]"""
    if x is not None and x.dtype != torch.float32:
        return x[:, :, :T].float().permute(1, 0, 2)
    return x


def recon_fwd(y, stats, gamma, beta, x, x_hat, loss_sums, T, G, loss_kind, rowsums=None):
    x = _target(x, T)
    with torch.no_grad():
        xh = _recon_xhat(y, gamma, beta, T, G)
        if x_hat is not None:
            x_hat.copy_(xh)
        if x is not None:
            d = (xh - x).double()
            loss_sums[0] = _loss_terms(loss_kind, d).sum()
            loss_sums[1] = (d * d).sum()


def recon_bwd(y, stats, gamma, beta, x, g_loss, g_mse, inv_numel, dxhat_ext, dy, dgamma, dbeta, dbias, T, G, loss_kind,
              rowsums=None):
    x = _target(x, T)
    with torch.enable_grad():
        yl = y.detach().float().clone().requires_grad_(True)
        gl = gamma.detach().clone().requires_grad_(True)
        bl = beta.detach().clone().requires_grad_(True)
        xh = _recon_xhat(yl, gl, bl, T, G)
        obj = 0
        if x is not None:
            d = xh - x
            if g_loss is not None:
                obj = obj + g_loss.reshape(()) * inv_numel * _loss_terms(loss_kind, d).sum()
            if g_mse is not None:
                obj = obj + g_mse.reshape(()) * inv_numel * (d * d).sum()
        if dxhat_ext is not None:
            obj = obj + (xh * dxhat_ext).sum()
        gy, gg, gb = torch.autograd.grad(obj, [yl, gl, bl])
    gy = gy.clone()
    gy[:, :, T:] = 0
    write_planes(dy, gy[:, :, :T], T)
    dgamma.copy_(gg)
    dbeta.copy_(gb)
    dbias.copy_(gy.sum(dim=(1, 2)))


# ---- static fields (T = 1): compact [C, B] forms (csrc/static_ops.cu) -----------------------------------------
def pack_static(x, xc, xt=None):
    B, N = x.shape[0], x.shape[1]
    v = x.reshape(B, N).t()
    xc.copy_(v.to(xc.dtype).reshape(xc.shape))
    if xt is not None:
        xt.copy_(v.reshape(xt.shape))


def rows_compact16(padded, out):
    out.copy_(padded[..., 0].reshape(out.shape))


def rows_expand_f32(compact, padded, accumulate=False):
    c = compact.reshape(padded.shape[:-1])
    if accumulate:
        padded[..., 0] += c
    else:
        padded.zero_()
        padded[..., 0] = c


def static_stats(y, stats, G):
    gn_stats(y.float()[:, :, None], stats, 1, G)


def static_recon_ws(N, B, G, device):
    return torch.empty(4 * N + 6 * B * G + 8, dtype=torch.float32, device=device)


def static_recon_fwd(y, stats, gamma, beta, x, loss_sums, ws, G, loss_kind, xhat_t=None):
    N, B = y.shape
    xh = torch.empty(B, N, 1, dtype=torch.float32, device=y.device) if xhat_t is not None else None
    recon_fwd(y.float()[:, :, None], stats, gamma, beta, x.float().t()[:, :, None].contiguous(), xh, loss_sums, 1, G, loss_kind)
    if xhat_t is not None:
        xhat_t.copy_(xh[:, :, 0].t())


def static_recon_bwd(y, stats, gamma, beta, x, g_loss, g_mse, inv_numel, ws, dy, dgamma, dbeta, dbias, G, loss_kind):
    N, B = y.shape
    recon_bwd(y.float()[:, :, None], stats, gamma, beta, x.float().t()[:, :, None].contiguous(), g_loss, g_mse, inv_numel, None,
              dy.view(1, N, B, 1), dgamma, dbeta, dbias, 1, G, loss_kind)


# ---- heads --------------------------------------------------------------------------------------
def _head(h, w, sigma, bias, T):
    C, B, Tp = h.shape
    flat = h[:, :, :T].permute(1, 0, 2).reshape(B, C * T)
    return F.linear(flat, w / sigma, bias)


def head_fwd(h, w_orig, sigma, bias, out, T):
    with torch.no_grad():
        out.copy_(_head(h, w_orig, sigma, bias, T))


def head_bwd(h, w_orig, sigma, dout, dwn, dbias, dh, dh_accumulate, T):
    C, B, Tp = h.shape
    with torch.enable_grad():
        hl = h.detach().clone().requires_grad_(True)
        wn = (w_orig.detach() / sigma).requires_grad_(True)
        flat = hl[:, :, :T].permute(1, 0, 2).reshape(B, C * T)
        o = F.linear(flat, wn)
        gh, gw = torch.autograd.grad(o, [hl, wn], dout)
    dwn.copy_(gw)
    dbias.copy_(dout.sum(0))
    if dh is not None:
        if dh_accumulate:
            dh[:, :, :T] += gh[:, :, :T]
        else:
            dh.copy_(gh)


def _latent(z, w, sigma, bias, D, T):
    B = z.shape[0]
    o = F.linear(z, w / sigma, bias).reshape(B, D, T)
    return o.permute(1, 0, 2)                          # [D,B,T]


def latent_fwd(z, w_orig, sigma, bias, out, T):
    _, D, B, Tp = out.shape
    with torch.no_grad():
        o = _latent(z, w_orig, sigma, bias, D, T)
    write_planes(out, o, T)


def latent_bwd(z, w_orig, sigma, dact, dwn, dbias, dz, T):
    D, B, Tp = dact.shape
    with torch.enable_grad():
        zl = z.detach().clone().requires_grad_(True)
        wn = (w_orig.detach() / sigma).requires_grad_(True)
        bl = torch.zeros(D * T, device=z.device).requires_grad_(True)
        o = _latent(zl, wn, torch.ones_like(sigma), bl, D, T)
        gz, gw, gb = torch.autograd.grad(o, [zl, wn, bl], dact[:, :, :T])
    dwn.copy_(gw)
    dbias.copy_(gb)
    if dz is not None:
        dz.copy_(gz)


# ---- reparameterisation + KL --------------------------------------------------------------------
def _reparam_main(last, eps):
    L = last.shape[1] // 2
    mu, lv = last[:, :L], torch.clamp(last[:, L:], -30, 30)
    std = torch.clamp(torch.exp(0.5 * lv), 1e-8, 10.0)
    z = mu + eps * std
    kl = torch.mean(0.5 * torch.sum(mu ** 2 + torch.exp(lv) - lv - 1, dim=1), dim=0)
    return z, kl


def reparam_main_fwd(last, eps, z, kl_out):
    with torch.no_grad():
        zz, kl = _reparam_main(last, eps)
    z.copy_(zz)
    kl_out.copy_(kl.reshape(1))


def reparam_main_bwd(last, eps, dz, dkl, dlast):
    with torch.enable_grad():
        ll = last.detach().clone().requires_grad_(True)
        z, kl = _reparam_main(ll, eps)
        obj = 0
        if dz is not None:
            obj = obj + (z * dz).sum()
        if dkl is not None:
            obj = obj + kl * dkl.reshape(())
        (g,) = torch.autograd.grad(obj, [ll])
    dlast.copy_(g)


def _kl2_reparam(cz, cxz, eps, h, std_scale, T):
    C = cz.shape[0] // 2
    mu, lv = cz[:C, :, :T], cz[C:, :, :T]
    dm, dl = cxz[:C, :, :T], cxz[C:, :, :T]
    lvc, dlc = torch.clamp(lv, -30, 30), torch.clamp(dl, -30, 30)
    var = torch.exp(lvc) + 1e-8
    integrand = torch.exp(dlc) / var + (mu - dm) ** 2 / var - dlc + lvc - 1
    lvt = torch.clamp(lv + dl, -30, 30)
    std = torch.clamp(torch.exp(0.5 * lvt) * std_scale, 1e-8, 10.0)
    z = (mu + dm) + eps.permute(1, 0, 2) * std
    zs = z + (h[:, :, :T] if h is not None else 0)
    return zs, integrand.sum()


def kl2_reparam_fwd(cz, cxz, eps, h, std_scale, zs_op, zs_f32, kl_sum, T):
    with torch.no_grad():
        zs, s = _kl2_reparam(cz, cxz, eps, h, std_scale, T)
    if zs_op is not None:
        write_planes(zs_op, zs, T)
    if zs_f32 is not None:
        zs_f32.zero_()
        zs_f32[:, :, :T] = zs
    kl_sum.copy_(s.double().reshape(1))


def kl2_reparam_bwd(cz, cxz, eps, std_scale, dzs, dkl, kl_scale, dcz, dcxz, T):
    with torch.enable_grad():
        a = cz.detach().clone().requires_grad_(True)
        b = cxz.detach().clone().requires_grad_(True)
        zs, s = _kl2_reparam(a, b, eps, None, std_scale, T)
        obj = 0
        if dzs is not None:
            obj = obj + (zs * dzs[:, :, :T]).sum()
        if dkl is not None:
            obj = obj + s * kl_scale * dkl.reshape(())
        ga, gb = torch.autograd.grad(obj, [a, b])
    for g in (ga, gb):
        g[:, :, T:] = 0
    dcz.copy_(ga)
    dcxz.copy_(gb)


# ---- RNG / optimiser ----------------------------------------------------------------------------
def counter_add(counter, inc):
    counter += inc


def philox_normal(out, seed, stream_id, sample0, counter=None):
    if counter is not None:
        stream_id = int(stream_id) + int(counter)
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(stream_id) * 7919 + int(sample0)) % (2 ** 63))
    out.copy_(torch.randn(out.shape, generator=g))


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq):
    gs = g * grad_scale
    if gnorm_sq is not None:
        gnorm_sq += (gs.double() ** 2).sum()
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(gs, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gs, gs, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)


def assemble_batch(data, ids, table, injected_noise, out, seed, draw, operand=None, blocks_per_sm=0):
    B = ids.shape[1]
    idx, other = ids[0].long(), ids[1].long()
    nl, sc, lam, om = table[0], table[1], table[2], table[3]
    s = data[idx]
    if injected_noise is not None:
        eps = injected_noise
    else:
        g = torch.Generator().manual_seed(int(seed) * 7919 + int(draw))
        eps = torch.randn(s.shape, generator=g).to(s.device)
    s = torch.where((nl != 0)[:, None, None], s + eps * nl[:, None, None], s)
    s = s * sc[:, None, None]
    mix = other >= 0
    partner = data[other.clamp_min(0)]
    s = torch.where(mix[:, None, None], lam[:, None, None] * s + om[:, None, None] * partner, s)
    if out is not None:
        out.copy_(s)
    if operand is not None:
        write_planes(operand, s.permute(1, 0, 2), data.shape[2])


def minmax_fit(data, rows, out_min, out_max, merge=False):
    import numpy as np
    x = data.cpu().numpy()
    if rows is not None:
        x = x[rows.cpu().numpy()]
    with np.errstate(all="ignore"):
        mn = np.fmin.reduce(x, axis=0, initial=np.inf) if x.shape[0] else np.full(x.shape[1], np.inf, x.dtype)
        mx = np.fmax.reduce(x, axis=0, initial=-np.inf) if x.shape[0] else np.full(x.shape[1], -np.inf, x.dtype)
    mn, mx = torch.from_numpy(mn.astype(x.dtype)).to(data.device), torch.from_numpy(mx.astype(x.dtype)).to(data.device)
    if merge:
        mn, mx = torch.fmin(out_min, mn), torch.fmax(out_max, mx)
    out_min.copy_(mn)
    out_max.copy_(mx)


def minmax_transform(data, scale, minv, out=None, out_t=None, T=0):
    val = data * scale          # two roundings, like X *= scale_; X += min_
    val = val + minv
    if out_t is not None:
        R, N = data.shape
        out_t.copy_(val.view(R // T, T, N).permute(0, 2, 1).to(torch.float32))
    if out is not None:
        out.copy_(val)


class OptPlan:
    """Emulated counterpart of kernels.OptPlan: keeps the item list (tensors) instead of a device table."""

    def __init__(self, items, device, dots=None, dot_base=0):
        self.items = items
        self.n = len(items)
        self.n_sn = sum(1 for it in items if it.get("u") is not None)
        self.dots = torch.zeros(self.n_sn + 6, dtype=torch.float64, device=device) if dots is None else dots
        d = dot_base
        for it in items:
            if it.get("u") is not None:
                it["dot_index"] = d
                d += 1


# ---- sharded optimiser over peer memory (csrc/optim.cu: peer_reduce_dot_kernel, opt_step_kernel with sg_peer) -----------
def make_peer(rank, weights, vecs, params, multicast=None):
    """Emulated sg_peer: the tensors themselves (every rank's weight-gradient arena, vector arena, flat parameters)."""
    return dict(rank=rank, world=len(weights), weights=list(weights), vecs=list(vecs), params=list(params))


def _same_region(buffers, mine, t):
    """the region tensor `t` occupies in buffers[mine], in every buffer"""
    off = (t.data_ptr() - buffers[mine].data_ptr()) // t.element_size()
    assert 0 <= off and off + t.numel() <= buffers[mine].numel()
    return [b[off:off + t.numel()].view(t.shape) for b in buffers]


def _w_gemm_layout(full):
    """weight_orig in the GEMM layout [k][Cout][Cin_p] of its gradient"""
    p, k, Cout, Cin, Cin_p = full["p"], full["k"], full["Cout"], full["Cin"], full["Cin_p"]
    w = torch.zeros(k, Cout, Cin_p, dtype=p.dtype, device=p.device)
    if full["flip"]:
        w[:, :, :Cin] = p.reshape(Cin, Cout, k).flip(2).permute(2, 1, 0)
    else:
        w[:, :, :Cin] = p.reshape(Cout, Cin, k).permute(2, 0, 1)
    return w


def peer_reduce_dot(plan, want_bad, peer, clear_dots=True, max_blocks=0):
    if clear_dots:
        plan.dots.zero_()
    me = peer["rank"]
    for it in plan.items:
        full, (lo, hi) = it["full"], it["rows"]
        arenas = peer["vecs"] if it.get("vec_arena") else peer["weights"]
        gs = _same_region(arenas, me, full["g"])
        total = sum(g.double() for g in gs).to(torch.float32)
        if full.get("u") is not None:
            k, Cout, Cin_p = full["k"], full["Cout"], full["Cin_p"]
            tot3, mine3 = total.view(k, Cout, Cin_p), gs[me].view(k, Cout, Cin_p)
            sel = (slice(None), slice(None), slice(lo, hi)) if full["flip"] else (slice(None), slice(lo, hi), slice(None))
            mine3[sel] = tot3[sel]
            w3 = _w_gemm_layout(full)
            dot = (tot3[sel].double() * w3[sel].double()).sum()
            plan.dots[it["dot_index"]] += dot
            bad = not bool(torch.isfinite(dot))
        else:
            gs[me].view(-1)[lo:hi] = total.view(-1)[lo:hi]
            bad = not bool(torch.isfinite(total.view(-1)[lo:hi]).all())
        if want_bad and bad:
            plan.dots[plan.dots.numel() - 6] += 1


_STEP_ARGS = {}      # dots.data_ptr() -> (skip, step, grad_scale) left by the last phase-0/2 call (the device copy of AdamArgs)


def opt_step(plan, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq, scaler=None, peer=None, phase=0,
             max_blocks=0):
    key = plan.dots.data_ptr()
    if phase == 3:           # update only, scalars of the preceding phase-2 call on the same dots buffer
        skip, step, grad_scale = _STEP_ARGS[key]
        if skip:
            return
        return _opt_step_sharded(plan, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq, peer)
    if scaler is not None:
        # sg_scaler_state semantics (csrc/optim.cu opt_prologue_kernel)
        fl = scaler[:5].view(torch.float32)
        used = float(fl[0])
        grad_scale = grad_scale / used
        if phase == 2:
            bad = float(plan.dots[plan.dots.numel() - 6]) != 0.0
        else:
            bad = any(not bool(torch.isfinite(it["g"]).all()) for it in plan.items)
        if bad:
            fl[0] = max(used * float(fl[2]), float(fl[3]))
            scaler[6] = 0
            scaler[8] += 1
            scaler[9] = 1
            gnorm_sq.fill_(float("inf"))
            _STEP_ARGS[key] = (True, step, grad_scale)
            return
        scaler[7] += 1
        scaler[9] = 0
        scaler[6] += 1
        if int(scaler[6]) >= int(scaler[5]):
            fl[0] = min(used * float(fl[1]), float(fl[4]))
            scaler[6] = 0
        step = int(scaler[7])
    _STEP_ARGS[key] = (False, step, grad_scale)
    if peer is not None or any("full" in it for it in plan.items):
        return _opt_step_sharded(plan, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq, peer)
    for it in plan.items:
        p = it["p"]
        if it.get("u") is not None:
            k, Cout, Cin, Cin_p = it["k"], it["Cout"], it["Cin"], it["Cin_p"]
            grad = torch.empty_like(p)
            sn_weight_grad(it["g"].reshape(k, Cout, Cin_p), p, it["u"], it["vv"], it["sigma"], grad, Cout, Cin, Cin_p, k,
                           0, 0, it["flip"])
        else:
            grad = it["g"].reshape(p.shape)
        adamw_step(p, grad, it["m"], it["v"], lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq)


def _opt_step_sharded(plan, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq, peer):
    """Shard items (trainer.shard_item): the gradient in this rank's arena already holds the sum over ranks on the shard's
    rows (peer_reduce_dot) and plan.dots the all-reduced <G, W>; update the shard's rows and store them to every rank."""
    me = peer["rank"] if peer is not None else 0
    for it in plan.items:
        full, (lo, hi) = it["full"], it["rows"]
        p = full["p"]
        if full.get("u") is not None:
            k, Cout, Cin, Cin_p, flip = full["k"], full["Cout"], full["Cin"], full["Cin_p"], full["flip"]
            G = full["g"].reshape(k, Cout, Cin_p)[:, :, :Cin]
            if flip:
                Gn = G.flip(0).permute(2, 1, 0)                   # [Cin, Cout, k]
                uv = torch.outer(full["u"], full["vv"]).reshape(Cout, Cin, k).permute(1, 0, 2)
            else:
                Gn = G.permute(1, 2, 0)                          # [Cout, Cin, k]
                uv = torch.outer(full["u"], full["vv"]).reshape(Cout, Cin, k)
            sigma = full["sigma"]
            dot = plan.dots[it["dot_index"]].float()
            grad = ((Gn - dot / sigma * uv) / sigma).reshape(p.shape)[lo:hi].reshape(-1)
            psh = p.reshape(p.shape[0], -1)[lo:hi].reshape(-1)
        else:
            grad = full["g"].reshape(-1)[lo:hi]
            psh = p.reshape(-1)[lo:hi]
        pnew = psh.clone()
        adamw_step(pnew, grad.contiguous(), it["m"], it["v"], lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq)
        targets = _same_region(peer["params"], me, p) if peer is not None else [p]
        for t in targets:
            if full.get("u") is not None:
                t.reshape(t.shape[0], -1)[lo:hi] = pnew.view(hi - lo, -1)
            else:
                t.reshape(-1)[lo:hi] = pnew


class SnPlan:
    def __init__(self, layers, device, dtype):
        self.layers = layers
        self.n = len(layers)


def sn_prepare(plan, training):
    for L in plan.layers:
        if L.get("u") is not None:
            sn_power_iter(L["w"], L["u"], L["v"], L["sigma"], L["H"], L["Cin"], L["k"], L["so"], L["si"], training)
        else:
            L["sigma"].fill_(1.0)
        if L.get("wg") is not None:
            sn_pack_weight(L["w"], L["sigma"], L["wg"], L["H"], L["Cin"], L["Cin_p"], L["k"], L["so"], L["si"], L["flip"])


NAMES = ["conv_fprop16", "pack_static", "rows_compact16", "rows_expand_f32", "static_stats", "static_recon_ws", "static_recon_fwd", "static_recon_bwd", "counter_add", "conv_out16_ok", "set_sm_limit", "make_peer", "peer_reduce_dot", "OptPlan", "opt_step", "SnPlan", "sn_prepare", "assemble_batch", "minmax_fit", "minmax_transform", "pack_input", "unpack_f32", "axpy", "scale_f64_to_f32", "sn_power_iter", "sn_pack_weight", "sn_weight_grad",
         "conv_fprop", "conv_fprop_gn", "conv_dgrad", "conv_wgrad", "gn_stats", "gn_act_fwd", "gn_act_bwd", "recon_fwd", "recon_bwd",
         "head_fwd", "head_bwd", "latent_fwd", "latent_bwd", "reparam_main_fwd", "reparam_main_bwd", "kl2_reparam_fwd",
         "kl2_reparam_bwd", "philox_normal", "adamw_step"]


class install:
    """Context manager used by the CPU wiring tests: route simulgen_vae_b200.kernels to this module."""

    def __enter__(self):
        import sys
        from simulgen_vae_b200 import kernels as K, engine
        me = sys.modules[__name__]
        self._saved = {n: getattr(K, n) for n in NAMES}
        self._check = engine._check_input
        for n in NAMES:
            setattr(K, n, getattr(me, n))
        engine._check_input = lambda *a, **k: None
        return self

    def __exit__(self, *exc):
        from simulgen_vae_b200 import kernels as K, engine
        for n, f in self._saved.items():
            setattr(K, n, f)
        engine._check_input = self._check
        return False
