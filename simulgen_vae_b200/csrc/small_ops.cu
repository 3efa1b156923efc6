// Spectral-norm power iteration and weight packing, linear heads, latent expansion,
// reparameterisation + KL, counter-based RNG and the fused AdamW / grad-norm step.
//
// Reference arithmetic replaced:
//   common.py:15-37 -> torch/nn/utils/spectral_norm.py:62-114 (power iteration, W / sigma, backward)
//   encoder.py:138-142,158-165 (xs_linear, last_x_linear), decoder.py:133,143 (Linear + Unflatten)
//   decoder.py:187-212,218-223 + losses.py:8-48 + VAE_network.py:103-105,113 (reparam, kl, kl_2)
//   train.py:92,156-168 (AdamW defaults, global gradient L2 norm)
#include "common.cuh"

namespace sg {

// =============================================================================================
// spectral norm
// =============================================================================================
// W_mat[o][q], q = i*k + j, lives at w[o*so + i*si + j].
__device__ __forceinline__ long long wm_addr(int o, int q, int k, long long so, long long si) {
    int i = q / k, j = q - i * k;
    return (long long)o * so + (long long)i * si + j;
}

// vraw[q] += sum_{o in chunk} W[o][q] * u[o]        grid: (ceil(Wd/256), ceil(H/rows))
__global__ void sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ vraw,
                               int H, int Wd, int k, long long so, long long si, int rows) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Wd) return;
    int o_lo = blockIdx.y * rows, o_hi = min(H, o_lo + rows);
    int i = q / k, j = q - i * k;
    const float* base = w + (long long)i * si + j;
    float acc = 0.f;
    for (int o = o_lo; o < o_hi; ++o) acc += __ldg(base + (long long)o * so) * __ldg(u + o);
    atomicAdd(&vraw[q], acc);
}

// out[0] = sum x^2 (single block)
__global__ void vec_sumsq_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i] * (double)x[i];
    double t = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = (float)t;
}

// uraw[o] += sum_{q in chunk} W[o][q] * vin[q] * scale      grid: (H, ceil(Wd/chunk)), block 256
// scale = 1 / max(sqrt(vsq[0]), eps) when vsq != NULL (training), else 1.
__global__ void sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ vin, const float* __restrict__ vsq,
                              float* __restrict__ uraw, int H, int Wd, int k, long long so, long long si, int chunk) {
    __shared__ double sh[32];
    int o = blockIdx.x;
    int q_lo = blockIdx.y * chunk, q_hi = min(Wd, q_lo + chunk);
    float acc = 0.f;
    for (int q = q_lo + threadIdx.x; q < q_hi; q += blockDim.x) acc += __ldg(w + wm_addr(o, q, k, so, si)) * __ldg(vin + q);
    double t = block_sum((double)acc, sh);
    if (threadIdx.x == 0) {
        float scale = 1.f;
        if (vsq != nullptr) scale = 1.f / fmaxf(sqrtf(vsq[0]), 1e-12f);
        atomicAdd(&uraw[o], (float)t * scale);
    }
}

// training: v = vraw / max(|vraw|, eps); u = uraw / max(|uraw|, eps); sigma = u . uraw
// eval    : sigma = u . uraw (u, v untouched)                              (single block)
__global__ void sn_finalize_kernel(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ uraw,
                                   const float* __restrict__ vraw, const float* __restrict__ vsq, float* __restrict__ sigma,
                                   int H, int Wd, int training) {
    __shared__ double sh[32];
    __shared__ float s_inv;
    if (training) {
        float inv_v = 1.f / fmaxf(sqrtf(vsq[0]), 1e-12f);
        for (int q = threadIdx.x; q < Wd; q += blockDim.x) v[q] = vraw[q] * inv_v;
        double s = 0.0;
        for (int o = threadIdx.x; o < H; o += blockDim.x) s += (double)uraw[o] * (double)uraw[o];
        double t = block_sum(s, sh);
        if (threadIdx.x == 0) {
            float nrm = fmaxf((float)sqrt(t), 1e-12f);
            s_inv = 1.f / nrm;
            sigma[0] = (float)(t / (double)nrm);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < H; o += blockDim.x) u[o] = uraw[o] * s_inv;
    } else {
        double s = 0.0;
        for (int o = threadIdx.x; o < H; o += blockDim.x) s += (double)uraw[o] * (double)u[o];
        double t = block_sum(s, sh);
        if (threadIdx.x == 0) sigma[0] = (float)t;
    }
}

// Wg[jj][o][i] = w(o,i,j) / sigma, jj = flip ? k-1-j : j ; zero for Cin <= i < Cin_p
template <typename OT>
__global__ void sn_pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ sigma, OT* __restrict__ wg,
                                      int Cout, int Cin, int Cin_p, int k, long long so, long long si, int flip) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)k * Cout * Cin_p;
    if (idx >= total) return;
    int i = (int)(idx % Cin_p);
    long long t = idx / Cin_p;
    int o = (int)(t % Cout);
    int jj = (int)(t / Cout);
    int j = flip ? k - 1 - jj : jj;
    float val = 0.f;
    if (i < Cin) val = __ldg(w + (long long)o * so + (long long)i * si + j) / sigma[0];
    from_f(wg[idx], val);
}

// dot[0] += sum G[jj][o][i] * w(o,i,j)
__global__ void sn_grad_dot_kernel(const float* __restrict__ g, const float* __restrict__ w, double* __restrict__ dot,
                                   int Cout, int Cin, int Cin_p, int k, long long so, long long si, int flip) {
    __shared__ double sh[32];
    long long total = (long long)k * Cout * Cin_p;
    double s = 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % Cin_p);
        if (i >= Cin) continue;
        long long t = idx / Cin_p;
        int o = (int)(t % Cout);
        int jj = (int)(t / Cout);
        int j = flip ? k - 1 - jj : jj;
        s += (double)g[idx] * (double)__ldg(w + (long long)o * so + (long long)i * si + j);
    }
    double t = block_sum(s, sh);
    if (threadIdx.x == 0) atomicAdd(dot, t);
}

// grad(o,i,j) = (G - (dot / sigma) * u[o] * v[i*k+j]) / sigma      (dot = <G, W_orig>, so <G,W_n> = dot/sigma)
__global__ void sn_grad_apply_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                                     const float* __restrict__ sigma, const double* __restrict__ dot,
                                     float* __restrict__ grad, int Cout, int Cin, int Cin_p, int k, long long so,
                                     long long si, int flip) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)k * Cout * Cin_p;
    if (idx >= total) return;
    int i = (int)(idx % Cin_p);
    if (i >= Cin) return;
    long long t = idx / Cin_p;
    int o = (int)(t % Cout);
    int jj = (int)(t / Cout);
    int j = flip ? k - 1 - jj : jj;
    float sg_ = sigma[0];
    float coef = (float)(dot[0] / (double)sg_);
    grad[(long long)o * so + (long long)i * si + j] = (g[idx] - coef * u[o] * v[i * k + j]) / sg_;
}

// =============================================================================================
// encoder heads: out[b][o] = (sum_{c,t} w[o][c*T+t] h[c][b][t]) / sigma + bias[o]
// =============================================================================================
__global__ void head_init_kernel(float* out, const float* bias, int B, int O) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * O) out[i] = bias[i % O];
}

constexpr int kHeadWarps = 8;
// grid: (ceil(C / kHeadWarps), B); warp w handles channel c = blockIdx.x*8 + w for sample b
__global__ void __launch_bounds__(kHeadWarps * 32)
head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ sigma,
                float* __restrict__ out, int C, int B, int T, int Tp, int O) {
    extern __shared__ float part[];  // [kHeadWarps][O]
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int c = blockIdx.x * kHeadWarps + warp, b = blockIdx.y;
    bool active = c < C;
    const float* hrow = h + ((long long)(active ? c : 0) * B + b) * Tp;
    for (int o = 0; o < O; ++o) {
        float acc = 0.f;
        if (active) {
            const float* wrow = w + (long long)o * C * T + (long long)c * T;
            for (int t = lane; t < T; t += 32) acc += __ldg(wrow + t) * hrow[t];
        }
        acc = warp_sum(acc);
        if (lane == 0) part[warp * O + o] = acc;
    }
    __syncthreads();
    float inv = 1.f / sigma[0];
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kHeadWarps; ++k) s += part[k * O + o];
        atomicAdd(&out[b * O + o], s * inv);
    }
}

// grid: (ceil(T/128), C), block 128.  dwn[o][c*T+t] = sum_b dout[b][o] h[c][b][t];
// dh[c][b][t] (+)= sum_o w[o][c*T+t] / sigma * dout[b][o]
template <int OMAX>
__global__ void __launch_bounds__(128)
head_bwd_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ sigma,
                const float* __restrict__ dout, float* __restrict__ dwn, float* __restrict__ dh, int dh_accumulate,
                int C, int B, int T, int Tp, int O, int Bc) {
    extern __shared__ float sdout[];  // [Bc][O]: the upstream gradient, staged in chunks of Bc samples
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    const bool active = t < T;
    const float inv = 1.f / sigma[0];
    float acc[OMAX], wv[OMAX];
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
        acc[o] = 0.f;
        wv[o] = (active && o < O) ? __ldg(w + (long long)o * C * T + (long long)c * T + t) * inv : 0.f;
    }
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int nb = min(Bc, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * O; i += blockDim.x) sdout[i] = dout[(long long)b0 * O + i];
        __syncthreads();
        if (!active) continue;
        for (int b = 0; b < nb; ++b) {
            long long hi = ((long long)c * B + b0 + b) * Tp + t;
            float hv = h[hi];
            float d = 0.f;
#pragma unroll
            for (int o = 0; o < OMAX; ++o) {
                if (o < O) {
                    float g = sdout[b * O + o];
                    acc[o] += g * hv;
                    d += wv[o] * g;
                }
            }
            if (dh != nullptr) dh[hi] = dh_accumulate ? dh[hi] + d : d;
        }
    }
    if (!active) return;
#pragma unroll
    for (int o = 0; o < OMAX; ++o)
        if (o < O) dwn[(long long)o * C * T + (long long)c * T + t] = acc[o];
}

__global__ void colsum_kernel(const float* __restrict__ dout, float* __restrict__ dbias, int B, int O) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dout[b * O + o];
    dbias[o] = s;
}

// =============================================================================================
// latent expansion (Linear(D, D*T) + Unflatten): out[d][b][t] = (w[d*T+t][:] . z[b][:]) / sigma + bias[d*T+t]
// =============================================================================================
// warp per (d, b) row; `planes` shifted copies of the row are written (the consumer is a k5 conv)
template <typename OT>
__global__ void __launch_bounds__(256)
latent_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ sigma,
                  const float* __restrict__ bias, OT* __restrict__ out, int planes, long long pstride, int D, int B,
                  int T, int Tp) {
    extern __shared__ float sg_rows[];
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);   // row = d * B + b
    if (row >= (long long)D * B) return;
    int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (Tp + 8);
    int d = (int)(row / B), b = (int)(row % B);
    srow_clear_halo(srow, Tp, lane);
    const float* zrow = z + (long long)b * D;
    float inv = 1.f / sigma[0];
    for (int t = lane; t < Tp; t += 32) {
        float val = 0.f;
        if (t < T) {
            const float* wrow = w + ((long long)d * T + t) * D;
            float acc = 0.f;
            for (int e = 0; e < D; ++e) acc += __ldg(wrow + e) * __ldg(zrow + e);
            val = acc * inv + bias[d * T + t];
        }
        srow[4 + t] = val;
    }
    __syncwarp();
    store_row_planes(out, row * Tp, planes, pstride, srow, T, Tp, lane);
}

// dwn[(d*T+t)][e] = sum_b dact[d][b][t] z[b][e] ; dbias[d*T+t] = sum_b dact[d][b][t]
__global__ void latent_bwd_w_kernel(const float* __restrict__ z, const float* __restrict__ dact, float* __restrict__ dwn,
                                    float* __restrict__ dbias, int D, int B, int T, int Tp) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)D * T * D;
    if (idx >= total) return;
    int e = (int)(idx % D);
    int row = (int)(idx / D);  // d*T + t
    int d = row / T, t = row - d * T;
    float acc = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
        float g = dact[((long long)d * B + b) * Tp + t];
        acc += g * __ldg(z + b * D + e);
        sb += g;
    }
    dwn[idx] = acc;
    if (e == 0) dbias[row] = sb;
}

// dz[b][e] = sum_{d,t} w[d*T+t][e] / sigma * dact[d][b][t]        grid: (D, B), block 256
__global__ void latent_bwd_z_kernel(const float* __restrict__ w, const float* __restrict__ sigma,
                                    const float* __restrict__ dact, float* __restrict__ dz, int D, int B, int T, int Tp) {
    __shared__ double sh[32];
    int e = blockIdx.x, b = blockIdx.y;
    float acc = 0.f;
    for (int i = threadIdx.x; i < D * T; i += blockDim.x) {
        int d = i / T, t = i - d * T;
        acc += __ldg(w + (long long)i * D + e) * dact[((long long)d * B + b) * Tp + t];
    }
    double tsum = block_sum((double)acc, sh);
    if (threadIdx.x == 0) dz[b * D + e] = (float)(tsum / (double)sigma[0]);
}

// =============================================================================================
// reparameterisation + KL
// =============================================================================================
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ bool inrange(float x, float lo, float hi) { return x >= lo && x <= hi; }

// single block
__global__ void reparam_main_fwd_kernel(const float* __restrict__ last, const float* __restrict__ eps,
                                        float* __restrict__ z, float* __restrict__ kl_out, int B, int L) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < B * L; i += blockDim.x) {
        int b = i / L, d = i - b * L;
        float mu = last[b * 2 * L + d];
        float lv = clampf(last[b * 2 * L + L + d], -30.f, 30.f);
        float std = clampf(expf(0.5f * lv), 1e-8f, 10.f);
        z[i] = mu + eps[i] * std;
        s += (double)(mu * mu + expf(lv) - lv - 1.f);
    }
    double t = block_sum(s, sh);
    if (threadIdx.x == 0) kl_out[0] = (float)(0.5 * t / (double)B);
}

__global__ void reparam_main_bwd_kernel(const float* __restrict__ last, const float* __restrict__ eps,
                                        const float* __restrict__ dz, const float* __restrict__ dkl,
                                        float* __restrict__ dlast, int B, int L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    int b = i / L, d = i - b * L;
    float mu = last[b * 2 * L + d];
    float lv_raw = last[b * 2 * L + L + d];
    float lv = clampf(lv_raw, -30.f, 30.f);
    float s = expf(0.5f * lv);
    float gz = dz ? dz[i] : 0.f;
    float gk = dkl ? dkl[0] / (float)B : 0.f;
    float dmu = gz + gk * mu;
    float dlv = gk * 0.5f * (expf(lv) - 1.f);
    if (inrange(s, 1e-8f, 10.f)) dlv += gz * eps[i] * 0.5f * s;
    if (!inrange(lv_raw, -30.f, 30.f)) dlv = 0.f;
    dlast[b * 2 * L + d] = dmu;
    dlast[b * 2 * L + L + d] = dlv;
}

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = 256;

// warp per (c, b) row
template <typename OT>
__global__ void __launch_bounds__(kThreads)
kl2_reparam_fwd_kernel(const float* __restrict__ cz, const float* __restrict__ cxz, const float* __restrict__ eps,
                       const float* __restrict__ h, float std_scale, OT* __restrict__ zs_op, int planes, long long pstride,
                       float* __restrict__ zs_f32, double* __restrict__ kl_sum, int C, int B, int T, int Tp) {
    __shared__ double sh[32];
    extern __shared__ float sg_rows[];
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (Tp + 8);
    const bool multi = planes > 1 && zs_op != nullptr;
    float ks = 0.f;
    if (row < (long long)C * B) {
        if (multi) srow_clear_halo(srow, Tp, lane);
        int c = (int)(row / B), b = (int)(row % B);
        long long off_mu = row * Tp, off_lv = ((long long)(C + c) * B + b) * Tp;
        const float* erow = eps + ((long long)b * C + c) * T;
        for (int seg = lane; seg < Tp / 8; seg += 32) {
            F8 mu = load8(cz + off_mu + seg * 8), lv = load8(cz + off_lv + seg * 8);
            F8 dm = load8(cxz + off_mu + seg * 8), dl = load8(cxz + off_lv + seg * 8);
            F8 hv = load8(h + off_mu + seg * 8);
            F8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int t = seg * 8 + i;
                float val = 0.f;
                if (t < T) {
                    float lvc = clampf(lv.v[i], -30.f, 30.f), dlc = clampf(dl.v[i], -30.f, 30.f);
                    float var = expf(lvc) + 1e-8f;
                    float diff = mu.v[i] - dm.v[i];
                    ks += expf(dlc) / var + diff * diff / var - dlc + lvc - 1.f;
                    float lvt = clampf(lv.v[i] + dl.v[i], -30.f, 30.f);
                    float std = clampf(expf(0.5f * lvt) * std_scale, 1e-8f, 10.f);
                    val = hv.v[i] + (mu.v[i] + dm.v[i]) + __ldg(erow + t) * std;
                }
                o.v[i] = val;
            }
            if (multi) {
#pragma unroll
                for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
            } else if (zs_op != nullptr) {
                store8(zs_op + off_mu + seg * 8, o);
            }
            if (zs_f32 != nullptr) store8(zs_f32 + off_mu + seg * 8, o);
        }
        if (multi) {
            __syncwarp();
            store_row_planes(zs_op, off_mu, planes, pstride, srow, T, Tp, lane);
        }
    }
    double t = block_sum((double)ks, sh);
    if (threadIdx.x == 0) atomicAdd(kl_sum, t);
}

template <typename OT>
__global__ void __launch_bounds__(kThreads)
kl2_reparam_bwd_kernel(const float* __restrict__ cz, const float* __restrict__ cxz, const float* __restrict__ eps,
                       float std_scale, const float* __restrict__ dzs, const float* __restrict__ dkl, float kl_scale,
                       OT* __restrict__ dcz, OT* __restrict__ dcxz, int C, int B, int T, int Tp) {
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= (long long)C * B) return;
    int lane = threadIdx.x & 31;
    int c = (int)(row / B), b = (int)(row % B);
    long long off_mu = row * Tp, off_lv = ((long long)(C + c) * B + b) * Tp;
    const float* erow = eps + ((long long)b * C + c) * T;
    float s = dkl ? dkl[0] * kl_scale : 0.f;
    for (int seg = lane; seg < Tp / 8; seg += 32) {
        F8 mu = load8(cz + off_mu + seg * 8), lv = load8(cz + off_lv + seg * 8);
        F8 dm = load8(cxz + off_mu + seg * 8), dl = load8(cxz + off_lv + seg * 8);
        F8 gz;
        if (dzs != nullptr) gz = load8(dzs + off_mu + seg * 8);
        F8 o_mu, o_lv, o_dm, o_dl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int t = seg * 8 + i;
            float g_mu = 0.f, g_lv = 0.f, g_dm = 0.f, g_dl = 0.f;
            if (t < T) {
                float g = dzs ? gz.v[i] : 0.f;
                float lvc = clampf(lv.v[i], -30.f, 30.f), dlc = clampf(dl.v[i], -30.f, 30.f);
                float elv = expf(lvc), edl = expf(dlc);
                float var = elv + 1e-8f;
                float diff = mu.v[i] - dm.v[i];
                // reparameterisation path
                float sum_lv = lv.v[i] + dl.v[i];
                float lvt = clampf(sum_lv, -30.f, 30.f);
                float std = expf(0.5f * lvt) * std_scale;
                float g_sum = 0.f;
                if (inrange(std, 1e-8f, 10.f) && inrange(sum_lv, -30.f, 30.f)) g_sum = g * __ldg(erow + t) * 0.5f * std;
                // kl_2 path
                float k_mu = 2.f * diff / var * s;
                float k_dl = inrange(dl.v[i], -30.f, 30.f) ? (edl / var - 1.f) * s : 0.f;
                float k_lv = inrange(lv.v[i], -30.f, 30.f) ? (1.f - (edl + diff * diff) * elv / (var * var)) * s : 0.f;
                g_mu = g + k_mu;
                g_dm = g - k_mu;
                g_lv = g_sum + k_lv;
                g_dl = g_sum + k_dl;
            }
            o_mu.v[i] = g_mu; o_lv.v[i] = g_lv; o_dm.v[i] = g_dm; o_dl.v[i] = g_dl;
        }
        store8(dcz + off_mu + seg * 8, o_mu);
        store8(dcz + off_lv + seg * 8, o_lv);
        store8(dcxz + off_mu + seg * 8, o_dm);
        store8(dcxz + off_lv + seg * 8, o_dl);
    }
}

// =============================================================================================
// Philox4x32-10 + Box-Muller
// =============================================================================================
// counter (may be NULL): device-resident draw counter added to stream_id - a CUDA-graph replay of the training step
// then draws fresh noise every time although its kernel arguments are frozen (sg_counter_add advances it inside the graph)
__global__ void philox_normal_kernel(float* __restrict__ out, int B, long long per_sample, uint64_t seed,
                                     uint64_t stream_id, long long sample0, const long long* __restrict__ counter) {
    if (counter != nullptr) stream_id += (uint64_t)counter[0];
    long long quads = (per_sample + 3) / 4;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= quads * B) return;
    int b = (int)(idx / quads);
    long long qd = idx - (long long)b * quads;
    uint64_t sample = (uint64_t)(sample0 + b);
    float n[4];
    philox_normal4(seed, stream_id, sample, (uint64_t)qd, n);
    float* dst = out + (long long)b * per_sample + qd * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (qd * 4 + i < per_sample) dst[i] = n[i];
}

// =============================================================================================
// AdamW over a flat arena + squared gradient norm (torch.optim.AdamW semantics, amsgrad=False)
// =============================================================================================
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float grad_scale, double* __restrict__ gnorm_sq) {
    __shared__ double sh[32];
    double ss = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        ss += (double)gi * (double)gi;
        float pi = p[i] * (1.f - lr * wd);
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
    if (gnorm_sq != nullptr) {
        double t = block_sum(ss, sh);
        if (threadIdx.x == 0) atomicAdd(gnorm_sq, t);
    }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_sn_power_iter(const float* w_orig, float* u, float* v, float* sigma, float* ws, int H, int Cin, int k,
                     long long so, long long si, int training, void* stream) {
    cudaStream_t st = as_stream(stream);
    int Wd = Cin * k;
    float* vraw = ws;
    float* uraw = ws + Wd;
    float* vsq = ws + Wd + H;
    cudaMemsetAsync(uraw, 0, sizeof(float) * H, st);
    int chunk = 4096;
    dim3 g3(H, (unsigned)cdiv(Wd, chunk));
    if (training) {
        cudaMemsetAsync(vraw, 0, sizeof(float) * Wd, st);
        int rows = 64;
        dim3 g1((unsigned)cdiv(Wd, 256), (unsigned)cdiv(H, rows));
        sn_wt_u_kernel<<<g1, 256, 0, st>>>(w_orig, u, vraw, H, Wd, k, so, si, rows);
        vec_sumsq_kernel<<<1, 1024, 0, st>>>(vraw, Wd, vsq);
        sn_w_v_kernel<<<g3, 256, 0, st>>>(w_orig, vraw, vsq, uraw, H, Wd, k, so, si, chunk);
    } else {
        sn_w_v_kernel<<<g3, 256, 0, st>>>(w_orig, v, nullptr, uraw, H, Wd, k, so, si, chunk);
    }
    sn_finalize_kernel<<<1, 1024, 0, st>>>(u, v, uraw, vraw, vsq, sigma, H, Wd, training);
    return check_launch("sn_power_iter");
}

int sg_sn_pack_weight(const float* w_orig, const float* sigma, void* wg, int Cout, int Cin, int Cin_p, int k,
                      long long so, long long si, int flip, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    long long total = (long long)k * Cout * Cin_p;
    int grid = (int)cdiv(total, 256);
    if (is_op16(dtype))
        sn_pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w_orig, sigma, (__nv_bfloat16*)wg, Cout, Cin, Cin_p, k, so, si, flip);
    else
        sn_pack_weight_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_orig, sigma, (float*)wg, Cout, Cin, Cin_p, k, so, si, flip);
    return check_launch("sn_pack_weight");
}

int sg_sn_weight_grad(const float* dwg, const float* w_orig, const float* u, const float* v, const float* sigma,
                      float* grad, double* ws, int Cout, int Cin, int Cin_p, int k, long long so, long long si,
                      int flip, void* stream) {
    cudaStream_t st = as_stream(stream);
    long long total = (long long)k * Cout * Cin_p;
    cudaMemsetAsync(ws, 0, sizeof(double), st);
    int g1 = (int)(cdiv(total, 256) < 148 * 8 ? cdiv(total, 256) : 148 * 8);
    sn_grad_dot_kernel<<<g1, 256, 0, st>>>(dwg, w_orig, ws, Cout, Cin, Cin_p, k, so, si, flip);
    sn_grad_apply_kernel<<<(int)cdiv(total, 256), 256, 0, st>>>(dwg, u, v, sigma, ws, grad, Cout, Cin, Cin_p, k, so, si, flip);
    return check_launch("sn_weight_grad");
}

int sg_head_fwd(const float* h, const float* w_orig, const float* sigma, const float* bias, float* out, int C, int B,
                int T, int Tp, int O, void* stream) {
    cudaStream_t st = as_stream(stream);
    head_init_kernel<<<(B * O + 255) / 256, 256, 0, st>>>(out, bias, B, O);
    dim3 grid((unsigned)cdiv(C, kHeadWarps), B);
    head_fwd_kernel<<<grid, kHeadWarps * 32, sizeof(float) * kHeadWarps * O, st>>>(h, w_orig, sigma, out, C, B, T, Tp, O);
    return check_launch("head_fwd");
}

int sg_head_bwd(const float* h, const float* w_orig, const float* sigma, const float* dout, float* dwn, float* dbias,
                float* dh, int dh_accumulate, int C, int B, int T, int Tp, int O, void* stream) {
    SG_REQUIRE(O <= 64, "head_bwd: O=%d > 64 unsupported", O);
    cudaStream_t st = as_stream(stream);
    dim3 grid((unsigned)cdiv(T, 128), C);
    // the upstream gradient [B][O] is staged in shared memory in chunks of Bc samples (large-batch static fields: B = 512)
    const int Bc = (int)((size_t)B * O * sizeof(float) <= 32 * 1024 ? B : (32 * 1024) / (O * sizeof(float)));
    size_t sm = sizeof(float) * Bc * O;
    if (O <= 8)
        head_bwd_kernel<8><<<grid, 128, sm, st>>>(h, w_orig, sigma, dout, dwn, dh, dh_accumulate, C, B, T, Tp, O, Bc);
    else
        head_bwd_kernel<64><<<grid, 128, sm, st>>>(h, w_orig, sigma, dout, dwn, dh, dh_accumulate, C, B, T, Tp, O, Bc);
    colsum_kernel<<<(O + 63) / 64, 64, 0, st>>>(dout, dbias, B, O);
    return check_launch("head_bwd");
}

int sg_latent_fwd(const float* z, const float* w_orig, const float* sigma, const float* bias, void* out, int planes,
                  long long plane_stride, int D, int B, int T, int Tp, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "latent_fwd: planes must be 1, 3 or 5");
    int grid = (int)cdiv((long long)D * B, 8);
    size_t sm = sizeof(float) * 8 * (Tp + 8);
    if (is_op16(dtype))
        latent_fwd_kernel<__nv_bfloat16><<<grid, 256, sm, as_stream(stream)>>>(z, w_orig, sigma, bias, (__nv_bfloat16*)out, planes, plane_stride, D, B, T, Tp);
    else
        latent_fwd_kernel<float><<<grid, 256, sm, as_stream(stream)>>>(z, w_orig, sigma, bias, (float*)out, planes, plane_stride, D, B, T, Tp);
    return check_launch("latent_fwd");
}

int sg_latent_bwd(const float* z, const float* w_orig, const float* sigma, const float* dact, float* dwn, float* dbias,
                  float* dz, int D, int B, int T, int Tp, void* stream) {
    cudaStream_t st = as_stream(stream);
    long long total = (long long)D * T * D;
    latent_bwd_w_kernel<<<(int)cdiv(total, 256), 256, 0, st>>>(z, dact, dwn, dbias, D, B, T, Tp);
    if (dz != nullptr) {
        dim3 grid(D, B);
        latent_bwd_z_kernel<<<grid, 256, 0, st>>>(w_orig, sigma, dact, dz, D, B, T, Tp);
    }
    return check_launch("latent_bwd");
}

int sg_reparam_main_fwd(const float* last, const float* eps, float* z, float* kl_out, int B, int L, void* stream) {
    reparam_main_fwd_kernel<<<1, 1024, 0, as_stream(stream)>>>(last, eps, z, kl_out, B, L);
    return check_launch("reparam_main_fwd");
}

int sg_reparam_main_bwd(const float* last, const float* eps, const float* dz, const float* dkl, float* dlast, int B,
                        int L, void* stream) {
    reparam_main_bwd_kernel<<<(B * L + 255) / 256, 256, 0, as_stream(stream)>>>(last, eps, dz, dkl, dlast, B, L);
    return check_launch("reparam_main_bwd");
}

int sg_kl2_reparam_fwd(const float* cz, const float* cxz, const float* eps, const float* h, float std_scale,
                       void* zs_op, int planes, long long plane_stride, float* zs_f32, double* kl_sum, int C, int B,
                       int T, int Tp, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "kl2_reparam_fwd: planes must be 1, 3 or 5");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(kl_sum, 0, sizeof(double), st);
    int grid = (int)cdiv((long long)C * B, kWarpsPerBlock);
    size_t sm = (planes > 1 && zs_op) ? sizeof(float) * kWarpsPerBlock * (Tp + 8) : 0;
    if (is_op16(dtype))
        kl2_reparam_fwd_kernel<__nv_bfloat16><<<grid, kThreads, sm, st>>>(cz, cxz, eps, h, std_scale, (__nv_bfloat16*)zs_op, planes, plane_stride, zs_f32, kl_sum, C, B, T, Tp);
    else
        kl2_reparam_fwd_kernel<float><<<grid, kThreads, sm, st>>>(cz, cxz, eps, h, std_scale, (float*)zs_op, planes, plane_stride, zs_f32, kl_sum, C, B, T, Tp);
    return check_launch("kl2_reparam_fwd");
}

int sg_kl2_reparam_bwd(const float* cz, const float* cxz, const float* eps, float std_scale, const float* dzs,
                       const float* dkl, float kl_scale, float* dcz, float* dcxz, int C, int B, int T, int Tp,
                       void* stream) {
    cudaStream_t st = as_stream(stream);
    int grid = (int)cdiv((long long)C * B, kWarpsPerBlock);
    kl2_reparam_bwd_kernel<float><<<grid, kThreads, 0, st>>>(cz, cxz, eps, std_scale, dzs, dkl, kl_scale, dcz, dcxz, C, B, T, Tp);
    return check_launch("kl2_reparam_bwd");
}

int sg_philox_normal(float* out, int B, long long per_sample, unsigned long long seed, unsigned long long stream_id,
                     long long sample0, const long long* counter, void* stream) {
    long long quads = (per_sample + 3) / 4;
    long long total = quads * B;
    if (total <= 0) return 0;
    philox_normal_kernel<<<(int)cdiv(total, 256), 256, 0, as_stream(stream)>>>(out, B, per_sample, seed, stream_id, sample0, counter);
    return check_launch("philox_normal");
}

__global__ void counter_add_kernel(long long* c, long long inc) { c[0] += inc; }

int sg_counter_add(long long* counter, long long inc, void* stream) {
    counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, inc);
    return check_launch("counter_add");
}

int sg_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, double* gnorm_sq, void* stream) {
    if (n <= 0) return 0;
    float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    int grid = (int)(cdiv(n, 256) < 148 * 8 ? cdiv(n, 256) : 148 * 8);
    adamw_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                      sqrtf(bc2), grad_scale, gnorm_sq);
    return check_launch("adamw_step");
}

}  // extern "C"
