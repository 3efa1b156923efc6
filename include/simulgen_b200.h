/*
 * simulgen_b200.h - C ABI of the B200-native SimulGen-VAE hot-path engine (libsimulgen_b200.so).
 *
 * The reference (leesihun/SimulGen-VAE) is pure Python/PyTorch and defines no FFI of its own
 * (SURVEY.md 2.1, 8b); the drop-in boundary is its nn.Module API (modules/VAE_network.py:60-121,
 * encoder.py:116-167, decoder.py:106-223, common.py:15-162, losses.py:8-48).  The Python overlay in
 * simulgen_vae_b200/overlay/modules keeps that API and calls the entry points below through ctypes;
 * each one replaces the ATen/cuDNN/cuBLAS kernels the named reference lines launch implicitly.
 *
 * Conventions
 *   - plain pointers and sizes only: no torch types.  All pointers are DEVICE pointers owned by the
 *     caller (torch-allocated tensors); kernels never allocate or free (workspaces are passed in).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - return value: 0 = ok, non-zero = error (sg_last_error() gives the text); the Python side raises
 *     RuntimeError, the reference's only error convention (VAE_network.py:119-121).
 *   - dtype: SG_BF16 (tcgen05 tensor-core path, bf16 operands + fp32 accumulation), SG_F16 (the same kernels with
 *     IEEE fp16 operands: 3 more mantissa bits at the same tensor-core rate; gradients need loss scaling) or SG_F32
 *     (fp32 validation mode, SIMT kernels, no tensor cores).  One 16-bit format per process at a time.
 *   - internal activation layout "CR": [C][B][Tp] with Tp = roundup(T + 2, 8); entries t >= T of
 *     every (c, b) row are zero (they are the conv "same" padding shared by neighbouring samples),
 *     R = B * Tp.  External layout (reference): [B][C][T] fp32 (SimulGen-VAE.py:281-283).
 *   - GEMM operands (activations / output gradients that feed a k-tap conv) are stored as `planes`
 *     (1, 3 or 5 >= k) pre-shifted copies: op[pl][c][b][t] = value[c][b][t + pl - planes/2], zero outside
 *     [0, T); `plane_stride` is the element distance between planes.  TMA moves 16-byte granules, so
 *     a tap cannot be a 1-element shift of a box along the contiguous axis; it selects a plane.
 *   - conv weights in GEMM layout "Wg": [k][Cout][Cin_p], Cin_p = roundup(Cin, 8), already divided
 *     by the spectral norm sigma; ConvTranspose1d weights are stored as the equivalent Conv1d
 *     (taps flipped, channels swapped; decoder.py:31).
 */
#ifndef SIMULGEN_B200_H_
#define SIMULGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SG_BF16 = 0, SG_F32 = 1, SG_F16 = 2 };
enum { SG_ACT_NONE = 0, SG_ACT_GELU = 1, SG_ACT_TANH = 2 };
enum { SG_LOSS_MSE = 0, SG_LOSS_MAE = 1, SG_LOSS_SMOOTHL1 = 2, SG_LOSS_HUBER = 3 };

const char* sg_last_error(void);
int sg_version(void);
/* 1 if the tcgen05 path can run on the current device (sm_100), else 0. */
int sg_device_supported(void);

/* ---- layout ------------------------------------------------------------------------------- */
/* x[B][N][T] -> out[N][B][Tp] (dtype), zero gap.  Replaces the implicit layout the reference
 * feeds nn.Conv1d with (encoder.py:34, SimulGen-VAE.py:281-283).  x is fp32 (x_dtype = SG_F32, the reference's
 * format) or already in the 16-bit operand format (x_dtype == dtype, T % 8 == 0 = Tp: a staging buffer kept in 16 bits
 * so that the host -> device copy moves half the bytes; pure re-layout). */
int sg_pack_input(const void* x, int x_dtype, void* out, int B, int N, int T, int Tp, int dtype, void* stream);
/* out[B][C][T] fp32 <- in[C][B][Tp] fp32 (used for x_hat-style exports and tests). */
int sg_unpack_f32(const float* in, float* out, int B, int C, int T, int Tp, void* stream);
/* dst[i] (+)= alpha * src[i] (fp32), n elements. */
int sg_axpy_f32(float* dst, const float* src, float alpha, long long n, int accumulate, void* stream);
/* out(dtype)[i] = in(fp32)[i] */
int sg_cast_f32(const float* in, void* out, long long n, int dtype, void* stream);

/* ---- spectral norm (common.py:15-37 -> torch/nn/utils/spectral_norm.py:62-114) --------------
 * W_mat[o][q], q = i*k + j, is addressed inside w_orig as o*so + i*si + j
 * (Conv1d/Linear: so = Cin*k, si = k; ConvTranspose1d (dim=1): so = k, si = Cout*k).
 * training != 0: v <- normalize(W^T u), u <- normalize(W v) in place, sigma = u . (W v)
 * training == 0: sigma = u . (W v) with the stored vectors.  ws: >= (H + Wd + 4) floats. */
int sg_sn_power_iter(const float* w_orig, float* u, float* v, float* sigma, float* ws,
                     int H, int Cin, int k, long long so, long long si, int training, void* stream);
/* Wg[j'][o][i] = w(o,i,j) / sigma   (j' = j, or k-1-j if flip) ; rows padded to Cin_p with zeros. */
int sg_sn_pack_weight(const float* w_orig, const float* sigma, void* wg, int Cout, int Cin, int Cin_p, int k,
                      long long so, long long si, int flip, int dtype, void* stream);
/* Backward of W_n = W / sigma(W) with u, v constant (spectral_norm.py:97-113):
 *   dW = (G - <G, W_n> * u v^T) / sigma, where G = dWg (fp32, GEMM layout) and <G,W_n> = <G,W>/sigma.
 * ws: >= 2 doubles.  Writes grad in the w_orig layout (accumulate=0) . */
int sg_sn_weight_grad(const float* dwg, const float* w_orig, const float* u, const float* v, const float* sigma,
                      float* grad, double* ws, int Cout, int Cin, int Cin_p, int k, long long so, long long si,
                      int flip, void* stream);

/* ---- convolutions as implicit GEMM (encoder.py:34,43; common.py:84,110,135-141;
 *      decoder.py:31,118,135,145,155,164 and their autograd backward, train.py:153) --------------
 * fprop: out[Cout][R] (fp32) (+)= sum_{j,ci} Wg[j][co][ci] * act[ci][r + j - k/2] + bias[co]
 * dgrad: dx [Cin][R]  (fp32) (+)= sum_{j,co} Wg[j][co][ci] * dy[co][r - j + k/2]
 * wgrad: dWg[j][Cout][Cin_p] (fp32) = sum_r dy[co][r] * act[ci][r + j - k/2]
 * act / dy / Wg are `dtype` (bf16: tcgen05 + TMA kernels; fp32: SIMT validation kernels).
 * R % 8 == 0 and Cin_p % 8 == 0 are required (TMA global strides are multiples of 16 bytes).  */
int sg_conv_fprop(const void* wg, const void* act, int act_planes, long long act_plane_stride, const float* bias,
                  float* out, int Cin, int Cin_p, int Cout, int k, int R, int accumulate, int dtype, void* stream);
/* fprop with out [Cout][R] stored in the 16-bit operand format (no accumulate, no statistics; needs sg_conv_out16_ok(Cout)). */
int sg_conv_fprop16(const void* wg, const void* act, int act_planes, long long act_plane_stride, const float* bias,
                    void* out, int Cin, int Cin_p, int Cout, int k, int R, int dtype, void* stream);
/* fprop of a conv that feeds a GroupNorm: also returns stats[B][G][2] = (mean, rstd) of the output (bias included).
 * With the CTA-pair tensor-core kernel the sums are taken in the GEMM epilogue from the fp32 accumulators
 * (rowstat: >= 2*Cout*B floats of scratch) and the separate statistics pass over the output disappears; otherwise
 * the output is reduced by sg_gn_stats (ws: >= 2*B*G doubles).  out_bf16 != 0 (bf16 mode, Cout > 128): the output
 * is stored as bf16 [Cout][B][Tp]. */
int sg_conv_fprop_gn(const void* wg, const void* act, int act_planes, long long act_plane_stride, const float* bias,
                     void* out, int out_bf16, int Cin, int Cin_p, int Cout, int k, int B, int T, int Tp, int G,
                     float* stats, double* ws, float* rowstat, int dtype, void* stream);
/* dx [Cin][R]: fp32, or (dx_dtype = the 16-bit operand format; needs sg_conv_out16_ok(Cin)) 16-bit - the gradient of
 * an activation whose only consumer was this conv; accumulate then adds to the 16-bit content. */
int sg_conv_dgrad(const void* wg, const void* dy, int dy_planes, long long dy_plane_stride, void* dx, int dx_dtype,
                  int Cin, int Cin_p, int Cout, int k, int R, int accumulate, int dtype, void* stream);
/* 1 when a GEMM with M output rows (fprop: Cout, dgrad: Cin) runs on the CTA-pair kernel and can store 16 bits */
int sg_conv_out16_ok(int M);
/* Upper bound on the SMs the persistent GEMM grids of the following launches occupy (0 = all).  The data-parallel
 * trainer lowers it while NCCL collectives overlap the backward pass so that their kernels find free SMs. */
int sg_set_sm_limit(int sms);
int sg_conv_wgrad(const void* dy, int dy_planes, long long dy_plane_stride, const void* act, int act_planes,
                  long long act_plane_stride, float* dwg, int Cin, int Cin_p, int Cout, int k, int R, int dtype,
                  void* stream);

/* ---- GroupNorm + activation + residual (encoder.py:35-36, common.py:85-102, decoder.py:32,119-120)
 * stats[B][G][2] fp32 = (mean, 1/sqrt(biased variance + 1e-5)) over the group's (C/G) x T valid entries
 * (nn.GroupNorm semantics; accumulated and finalised in fp64).  ws: >= 2*B*G doubles. */
int sg_gn_stats(const float* y, double* ws, float* stats, int C, int B, int T, int Tp, int G, void* stream);
/* pre = res + res_scale * act(gamma * (y - mean) * rstd + beta)   (stats == NULL: no norm, y used as is)
 * out = post_gelu ? gelu(pre) : pre ; written as operand (out_op, dtype, gap zeroed) and/or fp32.
 * y [C][B][Tp] is fp32 or (y_dtype = the 16-bit operand format) the 16-bit pre-norm output that sg_conv_fprop_gn
 * stored with out_bf16 != 0 - its statistics were taken from the fp32 accumulators. */
int sg_gn_act_fwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta,
                  const void* res, int res_is_f32, float res_scale, int act, int post_gelu,
                  void* out_op, int planes, long long plane_stride, float* out_f32,
                  int C, int B, int T, int Tp, int G, int dtype, void* stream);
/* Backward of the above.  dout [C][B][Tp]: fp32, or (dout_dtype) the 16-bit format when the dgrad GEMM that produced it
 * stored 16 bits (sg_conv_dgrad with a 16-bit dx).  Writes dy (dtype, gap zeroed) = grad wrt y,
 * dgamma/dbeta (may be NULL when stats == NULL), dbias[C] = sum_{b,t} dy, and dres (fp32,
 * (+)= per dres_accumulate bit 0) when res != NULL.  ws: >= 2*B*G doubles.  dres_accumulate bit 1: dgamma, dbeta,
 * dbias and ws were zeroed by the caller (they are accumulated into with atomics).
 * GroupNorm layers (stats != NULL) run two passes and OVERWRITE dout with dz = dL/d(gamma*xhat+beta): pass 1 takes the
 * reductions and evaluates the activation derivative once, pass 2 is dy = rstd*(gamma*dz - m1 - xhat*m2).
 * 16-bit y / dout are accepted for GroupNorm layers only. */
int sg_gn_act_bwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta,
                  const void* res, int res_is_f32, float res_scale, int act, int post_gelu,
                  void* dout, int dout_dtype, void* dy, int planes, long long plane_stride,
                  float* dgamma, float* dbeta, float* dbias, float* dres, int dres_accumulate, double* ws,
                  int C, int B, int T, int Tp, int G, int dtype, void* stream);

/* ---- reconstruction head: Tanh(GroupNorm(y)) + losses (decoder.py:117-121, VAE_network.py:71-77,110-111)
 * y [N][B][Tp] fp32 or bf16 (y_dtype; bf16 only in bf16 mode: the pre-norm output of the recon conv is the largest
 * tensor of the step and is read three times); x, x_hat fp32 [B][N][T] (either may be NULL).  loss_sums[2] doubles (zeroed inside):
 * sum of the selected loss terms and sum of squared errors. */
/* rowsums (optional, fp32 [N*B][4], 16-byte aligned; needs x): per-(n,b)-row partial sums of the GroupNorm
 * backward reductions, taken while y and x are in registers anyway; sg_recon_bwd then needs one pass. */
/* x / x_dtype: the target, either fp32 [B][N][T] (SG_F32, the reference's layout) or the packed 16-bit operand of the
 * first encoder conv [N][B][Tp] (x_dtype = the operand format; T % 8 == 0, 16-bit y): the values the encoder consumed. */
int sg_recon_fwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta, const void* x,
                 int x_dtype, float* x_hat, double* loss_sums, float* rowsums, int N, int B, int T, int Tp, int G,
                 int loss_kind, void* stream);
/* dx_hat = g_loss[0]*inv_numel*loss'(x_hat-x) + g_mse[0]*inv_numel*2(x_hat-x) + dxhat_ext (each optional),
 * then backward through tanh and GroupNorm -> dy (dtype, 1 plane), dgamma, dbeta, dbias.
 * rowsums: the buffer sg_recon_fwd filled (or NULL: two passes; also used when dxhat_ext != NULL).
 * ws: >= 2*B*G + 2 doubles. */
int sg_recon_bwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta, const void* x,
                 int x_dtype, const float* g_loss, const float* g_mse, float inv_numel, const float* dxhat_ext,
                 const float* rowsums, void* dy, float* dgamma, float* dbeta, float* dbias, double* ws,
                 int N, int B, int T, int Tp, int G, int loss_kind, int dtype, void* stream);
/* ---- static fields (preset Dim2 = 1 -> num_time = 1, modules/utils.py:306; SimulGen-VAE.py:279-283 with [P, N, 1] fields;
 * csrc/static_ops.cu) -------------
 * Compact [C][B] forms (B % 8 == 0) of the two N-channel layers, encoder conv0 (encoder.py:34) and the reconstruction
 * head (decoder.py:117-121, VAE_network.py:110-111): to sg_conv_fprop / dgrad / wgrad a compact tensor is an activation
 * with B / 8 samples of 8 valid columns, so the GEMMs do no work on padding.
 * x fp32 [B][N] -> xc [N][B] (dtype) and, when xt != NULL, the fp32 transpose xt [N][B] (the loss target). */
int sg_pack_static(const float* x, void* xc, float* xt, int B, int N, int dtype, void* stream);
/* out[r] = in[r][0] of padded 16-bit rows [R][8];  out[r][0..8) = (in[r], 0 .. 0) fp32 (accumulate: out[r][0] += in[r]). */
int sg_rows_compact16(const void* in, void* out, long long R, void* stream);
int sg_rows_expand_f32(const float* in, float* out, long long R, int accumulate, void* stream);
/* stats[b][g] = (mean, rstd) of GroupNorm(G, N) over y [N][B] (y_dtype); ws: 2 * B * G doubles. */
int sg_static_stats(const void* y, int y_dtype, double* ws, float* stats, int N, int B, int G, void* stream);
/* Tanh(GroupNorm(y)) against x [N][B] (x_dtype: fp32 or the operand format): loss_sums[2] as sg_recon_fwd, plus the
 * reductions of the GroupNorm backward in ws (4 * N + 4 * B * G floats, followed by 2 * B * G + 2 floats of scratch for
 * sg_static_recon_bwd, which must see the same ws).  y: 16-bit operand format, B <= 2048.
 * xhat_t (optional): x_hat transposed, fp32 [N][B] (x_hat[b][n][0] = xhat_t[n][b]). */
int sg_static_recon_fwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta, const void* x,
                        int x_dtype, float* xhat_t, double* loss_sums, float* ws, int N, int B, int G, int loss_kind,
                        void* stream);
/* dy [N][B] (dtype, 16-bit), dgamma, dbeta, dbias [N] for upstream g_loss / g_mse (device scalars, either may be NULL). */
int sg_static_recon_bwd(const void* y, int y_dtype, const float* stats, const float* gamma, const float* beta, const void* x,
                        int x_dtype, const float* g_loss, const float* g_mse, float inv_numel, float* ws, void* dy,
                        float* dgamma, float* dbeta, float* dbias, int N, int B, int G, int loss_kind, int dtype, void* stream);
/* out[i] = (float)(in[i] * scale), n small. */
int sg_scale_f64_to_f32(const double* in, float* out, double scale, int n, void* stream);

/* ---- linear heads (encoder.py:138-142,158-165; decoder.py:133,143) --------------------------
 * head: out[b][o] = (sum_{c,t} w_orig[o][c*T+t] * h[c][b][t]) / sigma + bias[o];  h fp32 [C][B][Tp]. */
int sg_head_fwd(const float* h, const float* w_orig, const float* sigma, const float* bias, float* out,
                int C, int B, int T, int Tp, int O, void* stream);
/* dwn[o][c*T+t] = sum_b dout[b][o] h[c][b][t] (grad wrt the normalised weight), dbias[o] = sum_b dout[b][o],
 * dh[c][b][t] (+)= sum_o w_orig[o][c*T+t]/sigma * dout[b][o]. */
int sg_head_bwd(const float* h, const float* w_orig, const float* sigma, const float* dout, float* dwn,
                float* dbias, float* dh, int dh_accumulate, int C, int B, int T, int Tp, int O, void* stream);
/* latent: out[d][b][t] = (sum_e w_orig[d*T+t][e] * z[b][e]) / sigma + bias[d*T+t]   (Linear + Unflatten) */
int sg_latent_fwd(const float* z, const float* w_orig, const float* sigma, const float* bias, void* out,
                  int planes, long long plane_stride, int D, int B, int T, int Tp, int dtype, void* stream);
int sg_latent_bwd(const float* z, const float* w_orig, const float* sigma, const float* dact, float* dwn,
                  float* dbias, float* dz, int D, int B, int T, int Tp, void* stream);

/* ---- reparameterisation + KL (decoder.py:187-212,218-223; losses.py:8-48; VAE_network.py:103-105,113)
 * main latent: last[B][2L] = (mu | log_var); z = mu + eps * clamp(exp(.5 clamp(lv)),1e-8,10);
 * kl_out[0] = mean_b(.5 sum_d(mu^2 + e^lv - lv - 1)). */
int sg_reparam_main_fwd(const float* last, const float* eps, float* z, float* kl_out, int B, int L, void* stream);
int sg_reparam_main_bwd(const float* last, const float* eps, const float* dz, const float* dkl, float* dlast,
                        int B, int L, void* stream);
/* hierarchical level: cz, cxz fp32 [2C][B][Tp] = (mu | lv), (dmu | dlv); eps fp32 [B][C][T];
 * z = (mu+dmu) + eps*clamp(std_scale*exp(.5 clamp(lv+dlv)),1e-8,10); zs = h + z (h fp32 [C][B][Tp]);
 * kl_sum[0] (double, zeroed inside) = sum of the kl_2 integrand (caller scales by .5/B). */
int sg_kl2_reparam_fwd(const float* cz, const float* cxz, const float* eps, const float* h, float std_scale,
                       void* zs_op, int planes, long long plane_stride, float* zs_f32, double* kl_sum,
                       int C, int B, int T, int Tp, int dtype, void* stream);
/* dzs fp32 [C][B][Tp] (or NULL); dkl = d/d(kl_2 value) (device scalar or NULL), kl_scale = .5/B;
 * writes dcz, dcxz fp32 [2C][B][Tp] (gap zeroed) = gradients wrt the two condition-conv outputs. */
int sg_kl2_reparam_bwd(const float* cz, const float* cxz, const float* eps, float std_scale, const float* dzs,
                       const float* dkl, float kl_scale, float* dcz, float* dcxz, int C, int B, int T, int Tp,
                       void* stream);

/* ---- counter-based RNG (replaces torch.randn_like, decoder.py:221) ----------------------------
 * out[b][i] ~ N(0,1), keyed on (seed, stream_id, sample0 + b, i): identical for any batch split. */
int sg_philox_normal(float* out, int B, long long per_sample, unsigned long long seed, unsigned long long stream_id,
                     long long sample0, const long long* counter, void* stream);
/* counter (device, may be NULL) is added to stream_id on the device; sg_counter_add advances it.  A captured CUDA graph
 * of the training step keeps its kernel arguments, so the draw index has to live in device memory. */
int sg_counter_add(long long* counter, long long inc, void* stream);

/* ---- optimiser (train.py:92,156-168: AdamW defaults + global grad L2 norm) ----------------------
 * One launch over a flat fp32 parameter arena.  gnorm_sq (double, (+)=) receives sum g^2. */
int sg_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, double* gnorm_sq, void* stream);

/* ---- batch assembly + augmentation (modules/augmentation.py:43-124, utils.py:56-66) ----------
 * data fp32 [P][N][T] resident on the GPU.  ids int32 [2][B] = (dataset index, mixup partner or -1);
 * table fp32 [4][B] = (noise_level or 0, scale, lambda, 1 - lambda).  out fp32 [B][N][T] =
 *     lambda * scale * (data[idx] + noise_level * eps) + (1 - lambda) * data[partner]
 * (reference order: noise, scaling, mixup; every product and sum rounded separately like the ATen sequence).  eps: injected_noise [B][N][T] if given, else Philox normals keyed on
 * (seed, draw, dataset index, element).  operand (optional): bf16 [N][B][Tp] copy for the first encoder conv. */
/* out may be NULL (operand-only batch).  blocks_per_sm: resident thread blocks per SM (0 = default 16; a loader that
 * prefetches underneath the training step passes 2 so that the step's GEMM CTAs still fit on every SM). */
int sg_assemble_batch(const float* data, int P, const int* ids, const float* table, const float* injected_noise,
                      float* out, void* operand, int B, int N, int T, int Tp, unsigned long long seed,
                      unsigned long long draw, int blocks_per_sm, void* stream);

/* ---- batched spectral-norm preparation: every layer of a sub-network in five launches ---------------
 * Same arithmetic as sg_sn_power_iter + sg_sn_pack_weight per layer (spectral_norm.py:62-114): the power
 * iteration depends only on the weights and the stored u / v, so all layers are prepared up front.
 * ws: per-layer scratch of (Cin*k + H) floats inside one arena [ws_base, ws_base + ws_elems) that is
 * zeroed here.  has_sn == 0: sigma = 1 (layer not spectral-normalised).  has_wg == 0: no operand copy
 * (Linear layers: the head / latent kernels read w_orig and sigma directly). */
typedef struct {
    const float* w;      /* weight_orig */
    float* u;            /* weight_u [H] (updated when training) */
    float* v;            /* weight_v [Cin*k] (updated when training) */
    float* sigma;        /* [1] out */
    void* wg;            /* [k][H][Cin_p] out (dtype) or NULL */
    float* ws;           /* scratch */
    long long so, si;    /* W_mat[o][i*k+j] = w[o*so + i*si + j] */
    int H, Cin, k, Cin_p, flip, has_sn, has_wg, reserved;
} sg_sn_layer;
int sg_sn_prepare(const sg_sn_layer* layers_dev, const sg_sn_layer* layers_host, int n_layers, float* ws_base,
                  long long ws_elems, int training, int dtype, void* stream);

/* ---- multi-tensor optimiser step (train.py:92,156-168 for the whole model in two launches) -----------
 * One item per parameter tensor.  Plain tensors (u == NULL): g is the gradient in p's layout.
 * Spectral-normalised weights (u != NULL): g is the wgrad GEMM output dWg [k][Cout][Cin_p] (gradient wrt
 * W/sigma); the gradient wrt weight_orig, (G - (<G,W>/sigma) u v^T)/sigma (spectral_norm.py:97-113 backward),
 * is formed on the fly - `dot` (one double per item, inside `dots`) receives <G,W> in the first launch.
 * p is walked in its native layout: Conv1d/Linear [Cout][Cin][k] (flip == 0), ConvTranspose1d
 * [Cin][Cout][k] (flip == 1, taps reversed in G).  AdamW: torch.optim.AdamW semantics (amsgrad off);
 * gnorm_sq (+)= sum of squared (scaled) gradients.  items_dev / items_host: the same array on the device
 * (read by the kernels) and on the host (read to size the grid). */
typedef struct {
    float* p;            /* parameter (updated in place) */
    const float* g;      /* gradient, see above */
    float* m;            /* exp_avg */
    float* v;            /* exp_avg_sq */
    const float* u;      /* weight_u [Cout] or NULL */
    const float* vv;     /* weight_v [Cin*k] */
    const float* sigma;  /* [1] */
    double* dot;         /* [1], scratch inside `dots` */
    long long n;         /* elements of p */
    int Cout, Cin, Cin_p, k, flip, reserved;
} sg_opt_item;
/* Device-resident dynamic loss scaler of the fp16-operand mode (the reference has no mixed precision; semantics are
 * torch.cuda.amp.GradScaler's).  The trainer multiplies the loss by `scale` ON THE DEVICE before backward; sg_opt_step
 * divides the gradients by it, and when any gradient is non-finite it skips the whole update (parameters, moments and
 * `step` untouched), multiplies `scale` by `backoff` and counts the event; `growth_interval` consecutive clean steps
 * multiply it by `growth`.  `step` is the AdamW bias-correction counter (advanced only by applied steps).
 * scaler == NULL: no scaling, bias corrections from the host's `step` argument.
 * dots: n_dots doubles of scratch, n_dots >= (number of spectral-norm items) + 6. */
typedef struct sg_scaler_state {
    float scale, growth, backoff, min_scale, max_scale;
    int growth_interval, good_steps, step, skipped, last_skipped;
} sg_scaler_state;
/* Data parallel over NVLink peer memory (one process per GPU; no reference counterpart - the reference has no working
 * multi-GPU path, modules/utils.py:209-238).  Every rank keeps its gradient arenas and ONE flat parameter buffer in
 * symmetric memory, so that each rank holds device pointers to all of them:
 *   wbase[r] / vbase[r]  weight-gradient arena / vector-gradient arena of rank r (same layout on every rank)
 *   pbase[r]             flat parameter buffer of rank r (same layout on every rank)
 * The optimiser is sharded (rank r owns a contiguous range of every tensor; `items` describe only that range):
 *   sg_peer_reduce_dot   for the elements of its shard, rank `rank` LOADS the gradient from every rank's arena over
 *                        NVLink, sums, stores the sum in its own arena and accumulates its share of <G, W> per layer
 *                        (a fused reduce-scatter + dot: no separate collective, no SM time taken from the backward pass);
 *   (host: one tiny all-reduce of `dots` - the only collective of the step)
 *   sg_opt_step(peer, phase 2)  spectral-norm gradient + AdamW on the shard; the updated parameters are STORED to every
 *                        rank's parameter buffer over NVLink (a fused all-gather).
 * Optimiser HBM traffic and state per GPU shrink by the world size; the gradient is never all-reduced. */
#define SG_MAX_PEERS 8
typedef struct sg_peer {
    int world, rank;
    const char* wbase[SG_MAX_PEERS];
    const char* vbase[SG_MAX_PEERS];
    char* pbase[SG_MAX_PEERS];
    /* NVSwitch multicast (NVLS) addresses of the same three buffers, or NULL: one multimem.ld_reduce returns the sum
     * over ALL ranks' copies (reduced inside the switch: ingress 1/W of the P2P-load path), one multimem.st writes
     * every rank's copy (egress 1/W of the P2P-store path). */
    const char* wmc;
    const char* vmc;
    char* pmc;
} sg_peer;
int sg_peer_reduce_dot(const sg_opt_item* items_dev, const sg_opt_item* items_host, int n_items, double* dots, int n_dots,
                       int want_bad_flag, int clear_dots, int max_blocks, const sg_peer* peer, void* stream);
/* phase 0: everything (dot pass, scalars, update).  phase 2: the dot pass has been done (sg_peer_reduce_dot + the
 * all-reduce of dots): scalars + update only.  phase 3: update only, with the scalars a phase-2 call on the SAME dots
 * buffer left there (the optimiser of one model split over two item tables - decoder / encoder - that share `dots`, so
 * that the exchange of the decoder's share overlaps the encoder's backward and the next forward).
 * peer (may be NULL): replicate the parameter stores to every rank. */
int sg_opt_step(const sg_opt_item* items_dev, const sg_opt_item* items_host, int n_items, double* dots, int n_dots,
                float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                double* gnorm_sq, sg_scaler_state* scaler, const sg_peer* peer, int phase, int max_blocks, void* stream);
/* max_blocks (sg_peer_reduce_dot, sg_opt_step): upper bound on the thread blocks of the launch (0 = one per 8192-element
 * chunk); a launch that runs underneath the backward / forward pass on another stream passes 2 x SM count. */

/* ---- preprocessing scan (modules/data_preprocess.py:65-165, SimulGen-VAE.py:279-283) --------------
 * data: the field matrix [R = P*T][N] (nodes innermost), float64 (is_f64 = 1) or float32 - the dtype the reference's
 * MinMaxScaler would compute in.
 * sg_minmax_fit: out_min/out_max [N] <- nanmin / nanmax over the rows `rows[0..n_rows)` (device int64 indices; NULL =
 *   rows 0..n_rows-1), i.e. MinMaxScaler.fit(FOM_data_aug[param_indices, time_indices, :]) (data_preprocess.py:108-113)
 *   without materialising the sampled copy; merge = 1 folds the result into the existing out_min/out_max (chunked
 *   datasets, MinMaxScaler.partial_fit semantics).  ws: scratch of ws_elems elements of the data dtype, >= 2 * N.
 * sg_minmax_transform: X * scale + min with two roundings (scaler.transform: X *= scale_; X += min_,
 *   data_preprocess.py:130-133) written to `out` ([R][N], same dtype, may alias data, may be NULL) and/or to `out_t`,
 *   the float32 [P][N][T] layout training consumes (new_x_train.transpose((0,2,1)) -> np.float32,
 *   SimulGen-VAE.py:282-283; T = time steps per parameter set, R % T == 0; NULL = skip). */
int sg_minmax_fit(const void* data, int is_f64, const long long* rows, long long n_rows, long long N, void* ws,
                  long long ws_elems, void* out_min, void* out_max, int merge, void* stream);
int sg_minmax_transform(const void* data, int is_f64, long long R, long long N, const void* scale, const void* minv,
                        void* out, float* out_t, long long T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMULGEN_B200_H_ */
