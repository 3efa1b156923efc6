"""Tensor-level wrappers around the C ABI (include/simulgen_b200.h).

Every function takes torch CUDA tensors that the caller allocated, passes raw device pointers,
sizes and the current CUDA stream to libsimulgen_b200.so, and returns nothing (outputs are written
in place).  There is no fallback: non-CUDA tensors or a missing extension raise RuntimeError.

The functions are looked up through the module attribute at call time (`K.conv_fprop(...)`), which
is what lets the CPU unit tests substitute `tests/kernel_emulator.py` to check the host-side wiring
without a GPU; the product never does that.
"""
import torch

from . import _lib

SG_BF16, SG_F32, SG_F16 = 0, 1, 2
OP16 = (torch.bfloat16, torch.float16)
ACT_NONE, ACT_GELU, ACT_TANH = 0, 1, 2
LOSS_KINDS = {"MSE": 0, "MAE": 1, "smoothL1": 2, "Huber": 3}

LAUNCHES = 0          # number of C-ABI calls (each launches >= 1 of our kernels); bench.py reports it
PROFILE = None        # bench.py sets this to a list: (name, algorithmic_flops, start_event, end_event) per GEMM


PROFILE_BYTES = 0     # algorithmic operand + output bytes of the GEMMs timed into PROFILE (bench.py: roofline.traffic's yardstick)


def _timed(name, flops, fn, nbytes=0):
    global PROFILE_BYTES
    if PROFILE is None:
        fn()
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PROFILE.append((name, flops, e0, e1))
    PROFILE_BYTES += nbytes


def _nb(*tensors):
    """bytes of the given tensors (each read or written once: the algorithmic minimum of a GEMM)"""
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


_HALF = False         # library variant of the next call: set by _dt() whenever a 16-bit operand is seen


def _dt(t):
    global _HALF
    if t.dtype == torch.bfloat16:
        _HALF = False
        return SG_BF16
    if t.dtype == torch.float32:
        return SG_F32
    if t.dtype == torch.float16:
        _HALF = True
        return SG_F16
    raise RuntimeError("simulgen_b200: unsupported operand dtype %s" % t.dtype)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("simulgen_b200: CUDA tensor required (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("simulgen_b200: contiguous tensor required, got strides %s for shape %s"
                           % (t.stride(), tuple(t.shape)))
    return t.data_ptr()


def _planes(t):
    """(pointer, planes, plane_stride) of an operand tensor [P, C, B, Tp] whose planes are contiguous
    [C, B, Tp] blocks (the tensor itself may be a channel slice of a wider buffer)."""
    if t is None:
        return None, 1, 0
    if not t.is_cuda:
        raise RuntimeError("simulgen_b200: CUDA tensor required (there is no CPU fallback)")
    if t.dim() != 4 or not t[0].is_contiguous():
        raise RuntimeError("simulgen_b200: operand must be [planes, C, B, Tp] with contiguous planes, got %s / %s"
                           % (tuple(t.shape), t.stride()))
    return t.data_ptr(), t.shape[0], (t.stride(0) if t.shape[0] > 1 else t[0].numel())


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream      # ~0.3 us; torch.cuda.current_stream() builds a Stream object (~3 us)
except AttributeError:  # pragma: no cover
    _raw_stream = None


def _stream():
    """cudaStream_t of the current stream of the current device (every C-ABI call takes it)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


PROFILE_ALL = None    # scripts/profile_step.py: list of (c_abi_name, start_event, end_event) for every call


def _call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE_ALL is None:
        _lib.call(name, *args, half=_HALF)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args, half=_HALF)
    e1.record()
    PROFILE_ALL.append((name, e0, e1))


def _f32(t, name):
    if t is not None and t.dtype != torch.float32:
        raise RuntimeError("simulgen_b200: %s must be float32" % name)
    return t


# ---- layout -------------------------------------------------------------------------------------
def pack_input(x, out, T):
    """x [B, N, T] fp32 (or already the operand dtype: pure re-layout); out: operand [1, N, B, Tp]."""
    B, N, _ = x.shape
    Tp = out.shape[3]
    assert out.shape[0] == 1 and (x.dtype == torch.float32 or x.dtype == out.dtype)
    od = _dt(out)
    _call("sg_pack_input", _p(x), od if x.dtype != torch.float32 else SG_F32, _p(out), B, N, T, Tp, od, _stream())


def unpack_f32(inp, out, T):
    C, B, Tp = inp.shape
    _call("sg_unpack_f32", _p(_f32(inp, "inp")), _p(_f32(out, "out")), B, C, T, Tp, _stream())


def axpy(dst, src, alpha, accumulate):
    assert dst.numel() == src.numel()
    _call("sg_axpy_f32", _p(_f32(dst, "dst")), _p(_f32(src, "src")), float(alpha), dst.numel(), int(accumulate), _stream())


def scale_f64_to_f32(inp, out, scale):
    assert inp.dtype == torch.float64 and out.dtype == torch.float32
    _call("sg_scale_f64_to_f32", _p(inp), _p(out), float(scale), inp.numel(), _stream())


# ---- spectral norm ------------------------------------------------------------------------------
def sn_power_iter(w_orig, u, v, sigma, H, Cin, k, so, si, training):
    ws = torch.empty(H + Cin * k + 8, dtype=torch.float32, device=w_orig.device)
    _call("sg_sn_power_iter", _p(_f32(w_orig, "w")), _p(u), _p(v), _p(sigma), _p(ws), H, Cin, k, so, si,
          int(training), _stream())


def sn_pack_weight(w_orig, sigma, wg, Cout, Cin, Cin_p, k, so, si, flip):
    _call("sg_sn_pack_weight", _p(_f32(w_orig, "w")), _p(sigma), _p(wg), Cout, Cin, Cin_p, k, so, si, int(flip),
          _dt(wg), _stream())


def sn_weight_grad(dwg, w_orig, u, v, sigma, grad, Cout, Cin, Cin_p, k, so, si, flip):
    ws = torch.empty(2, dtype=torch.float64, device=w_orig.device)
    _call("sg_sn_weight_grad", _p(_f32(dwg, "dwg")), _p(w_orig), _p(u), _p(v), _p(sigma), _p(_f32(grad, "grad")), _p(ws),
          Cout, Cin, Cin_p, k, so, si, int(flip), _stream())


# ---- convolutions -------------------------------------------------------------------------------
def conv_fprop(wg, act, bias, out, Cin, accumulate=False):
    """act: operand [P, Cin, B, Tp] (P >= k planes); out fp32 [Cout, B, Tp]."""
    k, Cout, Cin_p = wg.shape
    ap, an, astr = _planes(act)
    R = act.shape[2] * act.shape[3]
    assert act.shape[1] == Cin and out.shape[0] == Cout and out.numel() == Cout * R and wg.dtype == act.dtype
    _timed("fprop", 2.0 * Cin * Cout * k * R, lambda: _call(
        "sg_conv_fprop", _p(wg), ap, an, astr, _p(bias), _p(_f32(out, "out")), Cin, Cin_p, Cout, k, R, int(accumulate),
        _dt(act), _stream()), _nb(wg, act[:k], out))


def conv_fprop16(wg, act, bias, out, Cin):
    """fprop with out [Cout, B, Tp] in the operand dtype (no statistics): conv_out16_ok(Cout) layers only."""
    k, Cout, Cin_p = wg.shape
    ap, an, astr = _planes(act)
    R = act.shape[2] * act.shape[3]
    assert act.shape[1] == Cin and out.shape[0] == Cout and out.numel() == Cout * R and wg.dtype == act.dtype == out.dtype
    _timed("fprop", 2.0 * Cin * Cout * k * R, lambda: _call(
        "sg_conv_fprop16", _p(wg), ap, an, astr, _p(bias), _p(out), Cin, Cin_p, Cout, k, R, _dt(act), _stream()),
        _nb(wg, act[:k], out))


def conv_fprop_gn(wg, act, bias, out, Cin, stats, T, G):
    """Conv + GroupNorm statistics: out [Cout, B, Tp] (fp32, or bf16 for the recon layer in bf16 mode),
    stats fp32 [B, G, 2] <- (mean, rstd).  The statistics come from the GEMM epilogue when the CTA-pair kernel runs."""
    k, Cout, Cin_p = wg.shape
    ap, an, astr = _planes(act)
    B, Tp = act.shape[2], act.shape[3]
    assert act.shape[1] == Cin and out.shape[0] == Cout and out.numel() == Cout * B * Tp and wg.dtype == act.dtype
    out_bf16 = int(out.dtype in OP16)
    ws = torch.empty(2 * B * G, dtype=torch.float64, device=out.device)
    rowstat = torch.empty(2 * Cout * B, dtype=torch.float32, device=out.device) if act.dtype in OP16 else None
    _timed("fprop", 2.0 * Cin * Cout * k * B * Tp, lambda: _call(
        "sg_conv_fprop_gn", _p(wg), ap, an, astr, _p(bias), _p(out), out_bf16, Cin, Cin_p, Cout, k, B, T, Tp, int(G),
        _p(_f32(stats, "stats")), _p(ws), _p(rowstat), _dt(act), _stream()), _nb(wg, act[:k], out))


def conv_dgrad(wg, dy, dx, Cin, accumulate=False):
    """dy: operand [P, Cout, B, Tp]; dx [Cin, B, Tp] fp32, or the operand dtype when conv_out16_ok(Cin)."""
    k, Cout, Cin_p = wg.shape
    dp, dn, dstr = _planes(dy)
    R = dy.shape[2] * dy.shape[3]
    assert dy.shape[1] == Cout and dx.shape[0] == Cin and dx.numel() == Cin * R and wg.dtype == dy.dtype
    assert dx.dtype == torch.float32 or dx.dtype == dy.dtype
    dxd = _dt(dx)
    _timed("dgrad", 2.0 * Cin * Cout * k * R, lambda: _call(
        "sg_conv_dgrad", _p(wg), dp, dn, dstr, _p(dx), dxd, Cin, Cin_p, Cout, k, R, int(accumulate), _dt(dy),
        _stream()), _nb(wg, dy[:k], dx))


def set_sm_limit(sms):
    """SMs the persistent GEMM grids may use from now on (0 = all), in both library variants."""
    for half in (False, True):
        try:
            _lib.call("sg_set_sm_limit", int(sms), half=half)
        except RuntimeError:
            if not half:
                raise


_OUT16_OK = {}


def conv_out16_ok(M):
    """True when a GEMM with M output rows (fprop: Cout, dgrad: Cin) runs on the CTA-pair kernel, which can store its
    output in the 16-bit operand format."""
    r = _OUT16_OK.get(M)
    if r is None:
        r = _OUT16_OK[M] = bool(_lib.load(False).sg_conv_out16_ok(int(M)))
    return r


def conv_wgrad(dy, act, dwg, Cin):
    """dy: operand [Pd, Cout, B, Tp]; act: operand [Pa, Cin, B, Tp]; dwg fp32 [k, Cout, Cin_p]."""
    k, Cout, Cin_p = dwg.shape
    dp, dn, dstr = _planes(dy)
    ap, an, astr = _planes(act)
    R = dy.shape[2] * dy.shape[3]
    assert dy.shape[1] == Cout and act.shape[1] == Cin and dy.dtype == act.dtype
    _timed("wgrad", 2.0 * Cin * Cout * k * R, lambda: _call(
        "sg_conv_wgrad", dp, dn, dstr, ap, an, astr, _p(_f32(dwg, "dwg")), Cin, Cin_p, Cout, k, R, _dt(act), _stream()),
        _nb(dy[:1], act[:k], dwg))


# ---- GroupNorm + activation ---------------------------------------------------------------------
def gn_stats(y, stats, T, G):
    """stats: fp32 [B, G, 2] <- (mean, rstd) of every (sample, group)."""
    C, B, Tp = y.shape
    ws = torch.empty(2 * B * G, dtype=torch.float64, device=y.device)
    _call("sg_gn_stats", _p(_f32(y, "y")), _p(ws), _p(_f32(stats, "stats")), C, B, T, Tp, G, _stream())


def gn_act_fwd(y, stats, gamma, beta, res, res_scale, act, post_gelu, out_op, out_f32, T, G):
    """y: [C, B, Tp] fp32 or the 16-bit operand dtype (pre-norm conv output stored by conv_fprop_gn); out_op: operand
    [P, C, B, Tp] (all planes written) or None; res: [C, B, Tp] fp32 or operand dtype."""
    C, B, Tp = y.shape
    yd = _dt(y)
    dt = _dt(out_op) if out_op is not None else (yd if yd != SG_F32 else SG_F32)
    op, on, ostr = _planes(out_op)
    res_is_f32 = int(res is not None and res.dtype == torch.float32)
    _call("sg_gn_act_fwd", _p(y), yd, _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(res), res_is_f32, float(res_scale),
          int(act), int(post_gelu), op, on, ostr, _p(out_f32), C, B, T, Tp, int(G), dt, _stream())


def gn_act_bwd(y, stats, gamma, beta, res, res_scale, act, post_gelu, dout, dy, dgamma, dbeta, dbias, dres,
               dres_accumulate, T, G, ws=None):
    """ws: optional ZEROED float64 workspace of >= 2*B*G elements; passing it also declares dgamma / dbeta / dbias as
    already zeroed by the caller (the engine zeroes its whole gradient arena once per step).
    y / dout: fp32 or (GroupNorm layers) the 16-bit operand dtype.  GroupNorm layers OVERWRITE dout (with dz)."""
    C, B, Tp = y.shape
    yd, dd = _dt(y), _dt(dout)
    res_is_f32 = int(res is not None and res.dtype == torch.float32)
    flags = int(bool(dres_accumulate))
    if ws is None:
        ws = torch.empty(2 * B * max(int(G), 1) + 2, dtype=torch.float64, device=y.device)
    else:
        flags |= 2
    dp, dn, dstr = _planes(dy)
    _call("sg_gn_act_bwd", _p(y), yd, _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(res), res_is_f32, float(res_scale),
          int(act), int(post_gelu), _p(dout), dd, dp, dn, dstr, _p(dgamma), _p(dbeta), _p(dbias), _p(dres),
          flags, _p(ws), C, B, T, Tp, int(G), _dt(dy), _stream())


def recon_fwd(y, stats, gamma, beta, x, x_hat, loss_sums, T, G, loss_kind, rowsums=None):
    """rowsums: optional fp32 [N*B, 4] - partial sums of the GroupNorm backward taken by the forward.
    x: fp32 [B, N, T], or the packed 16-bit operand plane [N, B, Tp] of the encoder input (same dtype as y)."""
    N, B, Tp = y.shape
    xd = _dt(x) if x is not None else SG_F32
    assert x is None or (tuple(x.shape) == (B, N, T) if xd == SG_F32 else tuple(x.shape) == (N, B, Tp))
    _call("sg_recon_fwd", _p(y), _dt(y), _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(x), xd, _p(x_hat), _p(loss_sums),
          _p(_f32(rowsums, "rowsums")), N, B, T, Tp, G, int(loss_kind), _stream())


def recon_bwd(y, stats, gamma, beta, x, g_loss, g_mse, inv_numel, dxhat_ext, dy, dgamma, dbeta, dbias, T, G, loss_kind,
              rowsums=None):
    N, B, Tp = y.shape
    ws = torch.empty(2 * B * G + 2, dtype=torch.float64, device=y.device)
    dp, dn, _ = _planes(dy)
    assert dn == 1
    xd = _dt(x) if x is not None else SG_F32
    _call("sg_recon_bwd", _p(y), _dt(y), _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(x), xd, _p(g_loss), _p(g_mse), float(inv_numel),
          _p(dxhat_ext), _p(_f32(rowsums, "rowsums")), dp, _p(dgamma), _p(dbeta), _p(dbias), _p(ws), N, B, T, Tp, G,
          int(loss_kind), _dt(dy),
          _stream())


# ---- static fields (T = 1): compact [C, B] forms of the two N-channel layers (csrc/static_ops.cu) ---------------
def pack_static(x, xc, xt=None):
    """x fp32 [B, N, 1] -> xc [N, B] in the operand format (+ xt fp32 [N, B], the loss target); B % 8 == 0."""
    B, N = x.shape[0], x.shape[1]
    assert x.dtype == torch.float32 and x.numel() == B * N and xc.numel() == B * N and (xt is None or xt.numel() == B * N)
    _call("sg_pack_static", _p(x), _p(xc), _p(_f32(xt, "xt")), B, N, _dt(xc), _stream())


def rows_compact16(padded, out):
    """padded [..., 8] 16-bit rows -> out[...] = column 0."""
    assert padded.dtype in OP16 and out.dtype == padded.dtype and padded.shape[-1] == 8 and padded.numel() == 8 * out.numel()
    _dt(padded)
    _call("sg_rows_compact16", _p(padded), _p(out), out.numel(), _stream())


def rows_expand_f32(compact, padded, accumulate=False):
    """compact fp32 [...] -> padded fp32 [..., 8]: column 0 (+)= compact, columns 1..7 = 0 (untouched when accumulating)."""
    assert padded.shape[-1] == 8 and padded.numel() == 8 * compact.numel()
    _call("sg_rows_expand_f32", _p(_f32(compact, "compact")), _p(_f32(padded, "padded")), compact.numel(), int(accumulate), _stream())


def static_stats(y, stats, G):
    """y [N, B] (16-bit or fp32) -> stats fp32 [B, G, 2] = (mean, rstd) of GroupNorm(G, N) per sample."""
    N, B = y.shape
    ws = torch.empty(2 * B * G, dtype=torch.float64, device=y.device)
    _call("sg_static_stats", _p(y), _dt(y), _p(ws), _p(_f32(stats, "stats")), N, B, G, _stream())


def static_recon_ws(N, B, G, device):
    return torch.empty(4 * N + 4 * B * G + 2 * B * G + 8, dtype=torch.float32, device=device)


def static_recon_fwd(y, stats, gamma, beta, x, loss_sums, ws, G, loss_kind, xhat_t=None):
    """y, x [N, B] (y 16-bit; x fp32 or the operand format); ws = static_recon_ws(...), kept for static_recon_bwd;
    xhat_t: optional fp32 [N, B] <- x_hat transposed."""
    N, B = y.shape
    assert tuple(x.shape) == (N, B) and ws.numel() >= 4 * N + 6 * B * G + 2 and (xhat_t is None or tuple(xhat_t.shape) == (N, B))
    _call("sg_static_recon_fwd", _p(y), _dt(y), _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(x), _dt(x),
          _p(_f32(xhat_t, "xhat_t")), _p(loss_sums), _p(_f32(ws, "ws")), N, B, G, int(loss_kind), _stream())


def static_recon_bwd(y, stats, gamma, beta, x, g_loss, g_mse, inv_numel, ws, dy, dgamma, dbeta, dbias, G, loss_kind):
    N, B = y.shape
    assert tuple(dy.shape) == (N, B) and dy.dtype == y.dtype
    _call("sg_static_recon_bwd", _p(y), _dt(y), _p(_f32(stats, "stats")), _p(gamma), _p(beta), _p(x), _dt(x), _p(g_loss), _p(g_mse),
          float(inv_numel), _p(_f32(ws, "ws")), _p(dy), _p(dgamma), _p(dbeta), _p(dbias), N, B, G, int(loss_kind), _dt(dy), _stream())


# ---- linear heads -------------------------------------------------------------------------------
def head_fwd(h, w_orig, sigma, bias, out, T):
    C, B, Tp = h.shape
    O = w_orig.shape[0]
    _call("sg_head_fwd", _p(_f32(h, "h")), _p(w_orig), _p(sigma), _p(bias), _p(out), C, B, T, Tp, O, _stream())


def head_bwd(h, w_orig, sigma, dout, dwn, dbias, dh, dh_accumulate, T):
    C, B, Tp = h.shape
    O = w_orig.shape[0]
    _call("sg_head_bwd", _p(_f32(h, "h")), _p(w_orig), _p(sigma), _p(_f32(dout, "dout")), _p(dwn), _p(dbias), _p(dh),
          int(dh_accumulate), C, B, T, Tp, O, _stream())


def latent_fwd(z, w_orig, sigma, bias, out, T):
    """out: operand [P, D, B, Tp]."""
    _, D, B, Tp = out.shape
    op, on, ostr = _planes(out)
    _call("sg_latent_fwd", _p(_f32(z, "z")), _p(w_orig), _p(sigma), _p(bias), op, on, ostr, D, B, T, Tp, _dt(out),
          _stream())


def latent_bwd(z, w_orig, sigma, dact, dwn, dbias, dz, T):
    D, B, Tp = dact.shape
    _call("sg_latent_bwd", _p(_f32(z, "z")), _p(w_orig), _p(sigma), _p(_f32(dact, "dact")), _p(dwn), _p(dbias), _p(dz),
          D, B, T, Tp, _stream())


# ---- reparameterisation + KL --------------------------------------------------------------------
def reparam_main_fwd(last, eps, z, kl_out):
    B, L2 = last.shape
    _call("sg_reparam_main_fwd", _p(_f32(last, "last")), _p(_f32(eps, "eps")), _p(z), _p(kl_out), B, L2 // 2, _stream())


def reparam_main_bwd(last, eps, dz, dkl, dlast):
    B, L2 = last.shape
    _call("sg_reparam_main_bwd", _p(last), _p(eps), _p(dz), _p(dkl), _p(dlast), B, L2 // 2, _stream())


def kl2_reparam_fwd(cz, cxz, eps, h, std_scale, zs_op, zs_f32, kl_sum, T):
    C2, B, Tp = cz.shape
    dt = _dt(zs_op) if zs_op is not None else SG_F32
    zp, zn, zstr = _planes(zs_op)
    _call("sg_kl2_reparam_fwd", _p(_f32(cz, "cz")), _p(_f32(cxz, "cxz")), _p(_f32(eps, "eps")), _p(_f32(h, "h")),
          float(std_scale), zp, zn, zstr, _p(zs_f32), _p(kl_sum), C2 // 2, B, T, Tp, dt, _stream())


def kl2_reparam_bwd(cz, cxz, eps, std_scale, dzs, dkl, kl_scale, dcz, dcxz, T):
    C2, B, Tp = cz.shape
    _call("sg_kl2_reparam_bwd", _p(cz), _p(cxz), _p(eps), float(std_scale), _p(dzs), _p(dkl), float(kl_scale),
          _p(_f32(dcz, "dcz")), _p(_f32(dcxz, "dcxz")), C2 // 2, B, T, Tp, _stream())


# ---- RNG / optimiser ----------------------------------------------------------------------------
def philox_normal(out, seed, stream_id, sample0, counter=None):
    """counter: optional int64 [1] device tensor added to stream_id on the device (CUDA-graph replays)."""
    B = out.shape[0]
    per = out.numel() // B
    assert counter is None or (counter.dtype == torch.int64 and counter.numel() == 1)
    _call("sg_philox_normal", _p(_f32(out, "out")), B, per, int(seed) & (2 ** 64 - 1), int(stream_id), int(sample0),
          _p(counter), _stream())


def counter_add(counter, inc):
    assert counter.dtype == torch.int64 and counter.numel() == 1
    _call("sg_counter_add", _p(counter), int(inc), _stream())


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq):
    _call("sg_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
          float(weight_decay), int(step), float(grad_scale), _p(gnorm_sq), _stream())


class OptPlan:
    """Static description of one multi-tensor optimiser step (sg_opt_step): a list of items
    dict(p, g, m, v[, u, vv, sigma, Cout, Cin, Cin_p, k, flip]) whose tensors keep their addresses
    across steps (persistent gradient arena, optimiser state, spectral-norm buffers)."""

    def __init__(self, items, device, dots=None, dot_base=0):
        """dots / dot_base: share one scratch buffer between several plans (their <G, W> slots start at dot_base; the
        trailing 6 elements - overflow flag, step scalars - are common)."""
        import ctypes
        self.items = items
        n_sn = sum(1 for it in items if it.get("u") is not None)
        # one <G, W> per spectral-norm item + 6 doubles of scratch (non-finite flag, device copy of the step's scalars)
        self.dots = torch.zeros(n_sn + 6, dtype=torch.float64, device=device) if dots is None else dots
        assert self.dots.numel() >= dot_base + n_sn + 6
        arr = (_lib.OptItem * len(items))()
        d = dot_base
        for a, it in zip(arr, items):
            # sharded optimiser (shard_item): p / g / u / vv are VIEWS that start at the shard, "n" is the shard length
            a.p, a.g, a.m, a.v = it["p"].data_ptr(), it["g"].data_ptr(), _p(it["m"]), _p(it["v"])
            a.n = int(it.get("n", it["p"].numel()))
            a.reserved = int(bool(it.get("vec_arena", False)))
            assert it["m"].numel() == a.n and it["v"].numel() == a.n
            if it.get("u") is not None:
                a.u, a.vv, a.sigma = it["u"].data_ptr(), it["vv"].data_ptr(), _p(it["sigma"])
                a.dot = self.dots.data_ptr() + 8 * d
                it["dot_index"] = d
                d += 1
                a.Cout, a.Cin, a.Cin_p, a.k, a.flip = it["Cout"], it["Cin"], it["Cin_p"], it["k"], int(it["flip"])
                assert "n" in it or it["g"].numel() == it["k"] * it["Cout"] * it["Cin_p"]
            else:
                assert "n" in it or it["g"].numel() == a.n
        self._host = arr                                  # keeps the host copy alive (read at every launch)
        raw = bytes(arr)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.n = len(items)
        self.n_sn = n_sn


def make_peer(rank, weight_ptrs, vec_ptrs, param_ptrs, multicast=None):
    """sg_peer: device pointers to every rank's weight-gradient arena, vector-gradient arena and flat parameter buffer;
    multicast: optional (weights, vecs, params) NVSwitch multicast addresses of the same buffers."""
    world = len(weight_ptrs)
    assert 1 <= world <= 8 and len(vec_ptrs) == world and len(param_ptrs) == world and 0 <= rank < world
    pc = _lib.Peer()
    pc.world, pc.rank = world, rank

    def ptr(x):                 # a device pointer, or a tensor that starts at it (single-device tests of the kernels)
        return x.data_ptr() if isinstance(x, torch.Tensor) else int(x)
    for r in range(world):
        pc.wbase[r], pc.vbase[r], pc.pbase[r] = ptr(weight_ptrs[r]), ptr(vec_ptrs[r]), ptr(param_ptrs[r])
    if multicast is not None and all(multicast):
        pc.wmc, pc.vmc, pc.pmc = int(multicast[0]), int(multicast[1]), int(multicast[2])
    return pc


def peer_reduce_dot(plan, want_bad, peer, clear_dots=True, max_blocks=0):
    """Fused reduce-scatter + <G, W> over peer memory for the shard `plan` describes (csrc/optim.cu)."""
    import ctypes
    _call("sg_peer_reduce_dot", plan.table.data_ptr(), ctypes.addressof(plan._host), plan.n, plan.dots.data_ptr(),
          plan.dots.numel(), int(bool(want_bad)), int(bool(clear_dots)), int(max_blocks), ctypes.addressof(peer), _stream())


SCALER_FIELDS = ("scale", "growth", "backoff", "min_scale", "max_scale",          # float32
                 "growth_interval", "good_steps", "step", "skipped", "last_skipped")  # int32


def make_scaler_state(device, scale, growth=2.0, backoff=0.5, growth_interval=200, min_scale=2.0 ** -24, max_scale=2.0 ** 24):
    """Device copy of sg_scaler_state (include/simulgen_b200.h) as an int32[10] tensor whose first five words hold
    float32 bits.  `state[:1].view(torch.float32)` is the live loss scale."""
    import struct
    words = list(struct.unpack("5i", struct.pack("5f", float(scale), float(growth), float(backoff), float(min_scale),
                                                 float(max_scale)))) + [int(growth_interval), 0, 0, 0, 0]
    return torch.tensor(words, dtype=torch.int32, device=device)


def read_scaler_state(state):
    """Host copy as a dict (synchronises; logging and tests only)."""
    host = state.detach().cpu()
    f = host[:5].view(torch.float32).tolist()
    return dict(zip(SCALER_FIELDS, f + host[5:].tolist()))


def opt_step(plan, lr, beta1, beta2, eps, weight_decay, step, grad_scale, gnorm_sq, scaler=None, peer=None, phase=0,
             max_blocks=0):
    """scaler: optional make_scaler_state() tensor - the dynamic loss scale of the fp16-operand mode; `step` is then
    ignored (the applied-step counter lives in the state) and grad_scale excludes the loss scale.
    peer / phase: data parallel over peer memory - phase 2 runs after peer_reduce_dot + the all-reduce of plan.dots and
    stores the updated parameters to every rank (make_peer)."""
    import ctypes
    if scaler is not None:
        assert scaler.dtype == torch.int32 and scaler.numel() == len(SCALER_FIELDS)
    _call("sg_opt_step", plan.table.data_ptr(), ctypes.addressof(plan._host), plan.n, plan.dots.data_ptr(),
          plan.dots.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
          float(grad_scale), _p(gnorm_sq), _p(scaler), ctypes.addressof(peer) if peer is not None else None, int(phase),
          int(max_blocks), _stream())


class SnPlan:
    """Static description of one batched spectral-norm preparation (sg_sn_prepare): a list of layers
    dict(w, u, v, sigma, wg, H, Cin, k, Cin_p, so, si, flip) over persistent tensors."""

    def __init__(self, layers, device, dtype):
        import ctypes
        self.layers = layers
        self.dtype = {torch.bfloat16: SG_BF16, torch.float16: SG_F16}.get(dtype, SG_F32)
        total = sum(L["Cin"] * L["k"] + L["H"] + 8 for L in layers if L.get("u") is not None)
        self.ws = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        arr = (_lib.SnLayer * len(layers))()
        off = 0
        for a, L in zip(arr, layers):
            a.w, a.sigma = _p(L["w"]), _p(L["sigma"])
            a.H, a.Cin, a.k, a.Cin_p, a.flip = L["H"], L["Cin"], L["k"], L["Cin_p"], int(L["flip"])
            a.so, a.si = L["so"], L["si"]
            if L.get("u") is not None:
                a.u, a.v, a.has_sn = _p(L["u"]), _p(L["v"]), 1
                a.ws = self.ws.data_ptr() + 4 * off
                off += (L["Cin"] * L["k"] + L["H"] + 7) // 4 * 4
            if L.get("wg") is not None:
                assert L["wg"].is_contiguous() and tuple(L["wg"].shape) == (L["k"], L["H"], L["Cin_p"])
                a.wg, a.has_wg = _p(L["wg"]), 1
        self._host = arr
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.n = len(layers)


def sn_prepare(plan, training):
    import ctypes
    global _HALF
    if plan.dtype != SG_F32:
        _HALF = plan.dtype == SG_F16         # the plan's operand format selects the library build
    _call("sg_sn_prepare", plan.table.data_ptr(), ctypes.addressof(plan._host), plan.n, plan.ws.data_ptr(),
          plan.ws.numel(), int(training), plan.dtype, _stream())


# ---- batch assembly + augmentation --------------------------------------------------------------
def assemble_batch(data, ids, table, injected_noise, out, seed, draw, operand=None, blocks_per_sm=0):
    """data fp32 [P, N, T]; ids int32 [2, B]; table fp32 [4, B]; out fp32 [B, N, T] (None: operand only); operand:
    optional 16-bit [1, N, B, Tp] (the packed input of the first encoder conv, bf16 or fp16 like the engine's precision
    mode)."""
    P, N, T = data.shape
    B = ids.shape[1]
    assert out is None or tuple(out.shape) == (B, N, T)
    assert ids.dtype == torch.int32 and tuple(ids.shape) == (2, B) and tuple(table.shape) == (4, B)
    Tp = operand.shape[3] if operand is not None else 0
    if operand is not None:
        assert operand.dtype in OP16 and tuple(operand.shape[:3]) == (1, N, B)
        _dt(operand)                     # selects the library build for this 16-bit format
    _call("sg_assemble_batch", _p(_f32(data, "data")), P, _p(ids), _p(_f32(table, "table")), _p(_f32(injected_noise, "noise")),
          _p(_f32(out, "out")), _p(operand), B, N, T, Tp, int(seed) & (2 ** 64 - 1), int(draw), int(blocks_per_sm), _stream())


# ---- preprocessing scan ------------------------------------------------------------------
def _f64flag(t):
    if t.dtype == torch.float64:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise RuntimeError("simulgen_b200: the preprocessing scan works on float64 or float32 data, got %s" % t.dtype)


def minmax_fit(data, rows, out_min, out_max, merge=False):
    """data [R, N] float64/float32 (contiguous, nodes innermost); rows: int64 device tensor of row indices or None (all
    rows); out_min / out_max [N] (data dtype) <- nanmin / nanmax over those rows (merged into their content if merge)."""
    R, N = data.shape
    assert data.is_contiguous() and out_min.dtype == data.dtype and out_max.dtype == data.dtype
    assert out_min.numel() == N and out_max.numel() == N
    if rows is not None:
        assert rows.dtype == torch.int64 and rows.is_contiguous()
    n_rows = R if rows is None else rows.numel()
    ws = torch.empty(min(2 * N * 64, max(2 * N, 1 << 26)), dtype=data.dtype, device=data.device)
    _call("sg_minmax_fit", _p(data), _f64flag(data), _p(rows), n_rows, N, _p(ws), ws.numel(), _p(out_min), _p(out_max),
          int(merge), _stream())


def minmax_transform(data, scale, minv, out=None, out_t=None, T=0):
    """out [R, N] (may be `data` itself) <- data * scale + minv (two roundings); out_t fp32 [R // T, N, T] <- the same
    values cast to float32 in the [P, N, T] training layout."""
    R, N = data.shape
    assert data.is_contiguous() and scale.dtype == data.dtype and minv.dtype == data.dtype
    if out is not None:
        assert out.dtype == data.dtype and out.is_contiguous() and out.shape == data.shape
    if out_t is not None:
        assert out_t.dtype == torch.float32 and out_t.is_contiguous() and T > 0 and tuple(out_t.shape) == (R // T, N, T)
    _call("sg_minmax_transform", _p(data), _f64flag(data), R, N, _p(scale), _p(minv), _p(out), _p(out_t), int(T), _stream())
