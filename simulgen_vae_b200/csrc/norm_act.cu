// HBM-streaming kernels: layout packing, GroupNorm statistics / apply / backward, activation,
// residual, and the reconstruction head (tanh + fused loss reductions).
//
// Reference arithmetic replaced (all fp32 ATen kernels there):
//   encoder.py:35-36, common.py:85-102,108-125,133-162 (GroupNorm -> GELU -> x + 0.1 f(x)),
//   decoder.py:32 (GELU after ConvTranspose1d), decoder.py:117-121 (GroupNorm -> Tanh),
//   VAE_network.py:71-77,110-111 (MSE / L1 / SmoothL1 / Huber, mean reduction) and their backward.
//
// Every kernel is "one warp per (channel, sample) row": a row is Tp contiguous elements of which the
// first T are valid; each lane owns 8-element (16/32-byte) segments, so global accesses are
// coalesced 128-bit transactions; reductions use warp shuffles + one atomic per warp.
#include "common.cuh"

namespace sg {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

// ---------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(kThreads) pack_input_kernel(const float* __restrict__ x, OT* __restrict__ out,
                                                              int B, int N, int T, int Tp) {
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);  // row = n * B + b
    if (row >= (long long)N * B) return;
    int lane = threadIdx.x & 31;
    int n = (int)(row / B), b = (int)(row % B);
    const float* src = x + ((long long)b * N + n) * T;
    OT* dst = out + row * Tp;
    for (int seg = lane; seg < Tp / 8; seg += 32) {
        F8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int t = seg * 8 + i;
            r.v[i] = t < T ? __ldg(src + t) : 0.0f;
        }
        store8(dst + seg * 8, r);
    }
}

__global__ void __launch_bounds__(kThreads) unpack_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              int B, int C, int T, int Tp) {
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);  // row = c * B + b
    if (row >= (long long)C * B) return;
    int lane = threadIdx.x & 31;
    int c = (int)(row / B), b = (int)(row % B);
    const float* src = in + row * Tp;
    float* dst = out + ((long long)b * C + c) * T;
    for (int t = lane; t < T; t += 32) dst[t] = src[t];
}

__global__ void axpy_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, float alpha, long long n,
                                int accumulate) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = accumulate ? dst[i] + alpha * src[i] : alpha * src[i];
}

template <typename OT>
__global__ void cast_f32_kernel(const float* __restrict__ in, OT* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) from_f(out[i], in[i]);
}

__global__ void scale_f64_to_f32_kernel(const double* in, float* out, double scale, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(in[i] * scale);
}

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[b][g] += (sum, sumsq) over the group's rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const float* __restrict__ y, double* __restrict__ stats,
                                                            int C, int B, int T, int Tp, int G, int rows_per_block) {
    // grid: (chunks per group, G, B)
    __shared__ double sh[2][32];
    int g = blockIdx.y, b = blockIdx.z;
    int Cg = C / G;
    int c_lo = g * Cg + blockIdx.x * rows_per_block;
    int c_hi = min(c_lo + rows_per_block, (g + 1) * Cg);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s = 0.f, ss = 0.f;
    double ds = 0.0, dss = 0.0;
    for (int c = c_lo + warp; c < c_hi; c += kWarpsPerBlock) {
        const float* row = y + ((long long)c * B + b) * Tp;
        for (int seg = lane; seg < Tp / 8; seg += 32) {
            F8 r = load8(row + seg * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (seg * 8 + i < T) {
                    s += r.v[i];
                    ss += r.v[i] * r.v[i];
                }
            }
        }
        ds += (double)s;   // flush the fp32 partials per row to limit round-off growth
        dss += (double)ss;
        s = 0.f;
        ss = 0.f;
    }
    double t0 = block_sum(ds, sh[0]);
    double t1 = block_sum(dss, sh[1]);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[(size_t)(b * G + g) * 2], t0);
        atomicAdd(&stats[(size_t)(b * G + g) * 2 + 1], t1);
    }
}

// ---------------------------------------------------------------------------------------------
// forward apply
// ---------------------------------------------------------------------------------------------
template <typename OT, typename RT>
__global__ void __launch_bounds__(kThreads)
gn_act_fwd_kernel(const float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const RT* __restrict__ res, float res_scale, int act, int post_gelu,
                  OT* __restrict__ out_op, int planes, long long pstride, float* __restrict__ out_f32, int C, int B,
                  int T, int Tp, int G, double inv_n) {
    extern __shared__ float sg_rows[];
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= (long long)C * B) return;
    int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (Tp + 8);
    const bool multi = planes > 1 && out_op != nullptr;
    if (multi) srow_clear_halo(srow, Tp, lane);
    int c = (int)(row / B), b = (int)(row % B);
    float a = 1.f, sh = 0.f;
    if (stats != nullptr) {
        GnStat st = gn_stat(stats, b, c / (C / G), G, inv_n);
        a = gamma[c] * st.rstd;
        sh = beta[c] - st.mean * a;
    }
    const float* yrow = y + row * Tp;
    for (int seg = lane; seg < Tp / 8; seg += 32) {
        F8 yv = load8(yrow + seg * 8);
        F8 rv;
        if (res != nullptr) rv = load8(res + row * Tp + seg * 8);
        F8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float pre = res_scale * act_f(act, yv.v[i] * a + sh);
            if (res != nullptr) pre += rv.v[i];
            if (post_gelu) pre = gelu_f(pre);
            o.v[i] = (seg * 8 + i < T) ? pre : 0.f;
        }
        if (multi) {
#pragma unroll
            for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
        } else if (out_op != nullptr) {
            store8(out_op + row * Tp + seg * 8, o);
        }
        if (out_f32 != nullptr) store8(out_f32 + row * Tp + seg * 8, o);
    }
    if (multi) {
        __syncwarp();
        store_row_planes(out_op, row * Tp, planes, pstride, srow, T, Tp, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  The incoming gradient is either a tensor (generic layers) or derived from the
// reconstruction losses on the fly (recon head), so dx_hat is never materialised.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float loss_term(int kind, float d) {
    float ad = fabsf(d);
    if (kind == SG_LOSS_MAE) return ad;
    if (kind == SG_LOSS_SMOOTHL1 || kind == SG_LOSS_HUBER) return ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
    return d * d;
}
__device__ __forceinline__ float loss_grad(int kind, float d) {
    float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    if (kind == SG_LOSS_MAE) return sgn;
    if (kind == SG_LOSS_SMOOTHL1 || kind == SG_LOSS_HUBER) return fabsf(d) < 1.0f ? d : sgn;
    return 2.f * d;
}

struct BwdArgs {
    const float* y;
    const double* stats;
    const float* gamma;
    const float* beta;
    const void* res;
    float res_scale;
    int act, post_gelu;
    const float* dout;        // generic: [C][B][Tp]
    // loss-derived dout (recon head): x, ext in external layout [B][C][T]
    const float* x;
    const float* ext;
    float ga, gm;             // already multiplied by inv_numel
    int loss_kind;
    int C, B, T, Tp, G;
    double inv_n;
};

// Computes, for one 8-element segment, dyh = dL/d(gamma*xhat+beta) and xhat; returns dpre for dres.
template <typename RT, bool LOSS>
__device__ __forceinline__ void bwd_segment(const BwdArgs& p, long long row, int c, int b, int seg, float a, float sh,
                                            float mean, float rstd, F8& dyh, F8& xhat, F8& dpre) {
    F8 yv = load8(p.y + row * p.Tp + seg * 8);
    F8 rv, dv;
    const RT* res = reinterpret_cast<const RT*>(p.res);
    if (res != nullptr) rv = load8(res + row * p.Tp + seg * 8);
    if (!LOSS) dv = load8(p.dout + row * p.Tp + seg * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int t = seg * 8 + i;
        bool valid = t < p.T;
        float yh = yv.v[i] * a + sh;
        float val, dval;
        act_both(p.act, yh, val, dval);
        float d;
        if (LOSS) {
            d = 0.f;
            if (valid) {
                long long xi = ((long long)b * p.C + c) * p.T + t;
                if (p.x != nullptr) {
                    float diff = val - __ldg(p.x + xi);
                    d = p.ga * loss_grad(p.loss_kind, diff) + p.gm * 2.f * diff;
                }
                if (p.ext != nullptr) d += __ldg(p.ext + xi);
            }
        } else {
            d = dv.v[i];
        }
        float dp = d;
        if (p.post_gelu) {
            float pre = p.res_scale * val + (res != nullptr ? rv.v[i] : 0.f);
            dp = d * gelu_grad_f(pre);
        }
        float g = p.res_scale * dp * dval;
        dyh.v[i] = valid ? g : 0.f;
        xhat.v[i] = valid ? (yv.v[i] - mean) * rstd : 0.f;
        dpre.v[i] = valid ? dp : 0.f;
    }
}

template <typename RT, bool LOSS>
__global__ void __launch_bounds__(kThreads)
gn_bwd_reduce_kernel(BwdArgs p, float* __restrict__ dgamma, float* __restrict__ dbeta, double* __restrict__ S) {
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= (long long)p.C * p.B) return;
    int lane = threadIdx.x & 31;
    int c = (int)(row / p.B), b = (int)(row % p.B);
    int g = c / (p.C / p.G);
    GnStat st = gn_stat(p.stats, b, g, p.G, p.inv_n);
    float gm = p.gamma[c];
    float a = gm * st.rstd, sh = p.beta[c] - st.mean * a;
    float A = 0.f, Bx = 0.f;
    for (int seg = lane; seg < p.Tp / 8; seg += 32) {
        F8 dyh, xhat, dpre;
        bwd_segment<RT, LOSS>(p, row, c, b, seg, a, sh, st.mean, st.rstd, dyh, xhat, dpre);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            A += dyh.v[i];
            Bx += dyh.v[i] * xhat.v[i];
        }
    }
    A = warp_sum(A);
    Bx = warp_sum(Bx);
    if (lane == 0) {
        atomicAdd(&dgamma[c], Bx);
        atomicAdd(&dbeta[c], A);
        atomicAdd(&S[(size_t)(b * p.G + g) * 2], (double)(gm * A));
        atomicAdd(&S[(size_t)(b * p.G + g) * 2 + 1], (double)(gm * Bx));
    }
}

template <typename OT, typename RT, bool LOSS>
__global__ void __launch_bounds__(kThreads)
gn_bwd_apply_kernel(BwdArgs p, const double* __restrict__ S, OT* __restrict__ dy, int planes, long long pstride,
                    float* __restrict__ dbias, float* __restrict__ dres, int dres_accumulate) {
    extern __shared__ float sg_rows[];
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= (long long)p.C * p.B) return;
    int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (p.Tp + 8);
    const bool multi = planes > 1;
    if (multi) srow_clear_halo(srow, p.Tp, lane);
    int c = (int)(row / p.B), b = (int)(row % p.B);
    float a = 1.f, sh = 0.f, mean = 0.f, rstd = 1.f, gm = 1.f, m1 = 0.f, m2 = 0.f;
    bool has_gn = p.stats != nullptr;
    if (has_gn) {
        int g = c / (p.C / p.G);
        GnStat st = gn_stat(p.stats, b, g, p.G, p.inv_n);
        gm = p.gamma[c];
        mean = st.mean;
        rstd = st.rstd;
        a = gm * rstd;
        sh = p.beta[c] - mean * a;
        m1 = (float)(S[(size_t)(b * p.G + g) * 2] * p.inv_n);
        m2 = (float)(S[(size_t)(b * p.G + g) * 2 + 1] * p.inv_n);
    }
    float db = 0.f;
    for (int seg = lane; seg < p.Tp / 8; seg += 32) {
        F8 dyh, xhat, dpre, o;
        bwd_segment<RT, LOSS>(p, row, c, b, seg, a, sh, mean, rstd, dyh, xhat, dpre);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = dyh.v[i];
            if (has_gn) v = (seg * 8 + i < p.T) ? rstd * (gm * v - m1 - xhat.v[i] * m2) : 0.f;
            o.v[i] = v;
            db += v;
        }
        if (multi) {
#pragma unroll
            for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
        } else {
            store8(dy + row * p.Tp + seg * 8, o);
        }
        if (dres != nullptr) {
            float* dr = dres + row * p.Tp + seg * 8;
            if (dres_accumulate) {
                F8 old = load8(dr);
#pragma unroll
                for (int i = 0; i < 8; ++i) dpre.v[i] += old.v[i];
            }
            store8(dr, dpre);
        }
    }
    db = warp_sum(db);
    if (lane == 0 && dbias != nullptr) atomicAdd(&dbias[c], db);
    if (multi) {
        __syncwarp();
        store_row_planes(dy, row * p.Tp, planes, pstride, srow, p.T, p.Tp, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// recon head forward: x_hat = tanh(GN(y)) in the external layout + loss sums (+ the per-row partial
// sums of the GroupNorm backward, so that the backward needs ONE pass over y / x instead of two)
// ---------------------------------------------------------------------------------------------
// 8 consecutive elements of an external-layout row ([B][N][T], row start 16-byte aligned when VEC)
template <bool VEC>
__device__ __forceinline__ F8 load8_ext(const float* row, int t0, int T) {
    F8 r;
    if (VEC && t0 + 8 <= T) {
        float4 a = __ldg(reinterpret_cast<const float4*>(row + t0));
        float4 b = __ldg(reinterpret_cast<const float4*>(row + t0 + 4));
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = (t0 + i < T) ? __ldg(row + t0 + i) : 0.f;
    }
    return r;
}
template <bool VEC>
__device__ __forceinline__ void store8_ext(float* row, int t0, int T, const F8& r) {
    if (VEC && t0 + 8 <= T) {
        *reinterpret_cast<float4*>(row + t0) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
        *reinterpret_cast<float4*>(row + t0 + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (t0 + i < T) row[t0 + i] = r.v[i];
    }
}

// persistent: grid = min(rows / 8, a few waves); every warp strides over (n, b) rows.
// rowsums[row] = (sum gL, sum gL*xn, sum gM, sum gM*xn) with gL = loss'(d)(1 - xh^2), gM = 2 d (1 - xh^2),
// xn = (y - mean) rstd: the backward scales them by the upstream loss gradients (they are linear in them).
template <bool VEC, bool MSE>
__global__ void __launch_bounds__(kThreads)
recon_fwd_kernel(const float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ x, float* __restrict__ x_hat,
                 double* __restrict__ loss_sums, float4* __restrict__ rowsums, int N, int B, int T, int Tp, int G,
                 int loss_kind, double inv_n) {
    __shared__ double shm[2][32];
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)N * B;
    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    double d0 = 0.0, d1 = 0.0;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);
        GnStat st = gn_stat(stats, b, n / (N / G), G, inv_n);
        float a = gamma[n] * st.rstd, sh = beta[n] - st.mean * a;
        const float* yrow = y + row * Tp;
        long long xo = ((long long)b * N + n) * T;
        float l0 = 0.f, l1 = 0.f, aL = 0.f, bL = 0.f, aM = 0.f, bM = 0.f;
        for (int seg = lane; seg * 8 < T; seg += 32) {
            F8 yv = load8(yrow + seg * 8);
            F8 xv, xh;
            if (x != nullptr) xv = load8_ext<VEC>(x + xo, seg * 8, T);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool valid = seg * 8 + i < T;
                float h = tanh_fast(yv.v[i] * a + sh);
                xh.v[i] = h;
                if (x != nullptr && valid) {
                    float d = h - xv.v[i];
                    l1 += d * d;
                    if (!MSE) l0 += loss_term(loss_kind, d);
                    if (rowsums != nullptr) {
                        float om = 1.f - h * h;
                        float xn = (yv.v[i] - st.mean) * st.rstd;
                        float gm = 2.f * d * om;
                        aM += gm; bM += gm * xn;
                        if (!MSE) {
                            float gl = loss_grad(loss_kind, d) * om;
                            aL += gl; bL += gl * xn;
                        }
                    }
                }
            }
            if (x_hat != nullptr) store8_ext<VEC>(x_hat + xo, seg * 8, T, xh);
        }
        if (rowsums != nullptr) {
            aM = warp_sum(aM); bM = warp_sum(bM);
            if (MSE) { aL = aM; bL = bM; } else { aL = warp_sum(aL); bL = warp_sum(bL); }
            if (lane == 0) rowsums[row] = make_float4(aL, bL, aM, bM);
        }
        d0 += (double)(MSE ? l1 : l0);   // flush the fp32 partials per row
        d1 += (double)l1;
    }
    if (x != nullptr) {
        double t0 = block_sum(d0, shm[0]);
        double t1 = block_sum(d1, shm[1]);
        if (threadIdx.x == 0) {
            atomicAdd(&loss_sums[0], t0);
            atomicAdd(&loss_sums[1], t1);
        }
    }
}

// backward, step 1 of the one-pass path: fold the forward's row sums with the upstream scalars.
// grid (ceil(Cg / 64), G), block 256: warp w handles channels c0 + w, c0 + w + 8, ... (< 64 per block)
// dgamma[c] = sum_b Bx, dbeta[c] = sum_b A, S[b][g] += gamma_c * (A, Bx)  with (A, Bx) = ga*(.L) + gm*(.M)
__global__ void __launch_bounds__(kThreads)
recon_bwd_combine_kernel(const float4* __restrict__ rowsums, const float* __restrict__ scal,
                         const float* __restrict__ gamma, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         double* __restrict__ S, int N, int B, int G) {
    extern __shared__ float sacc[];   // [B][2]
    const int Cg = N / G, g = blockIdx.y;
    const int c_lo = g * Cg + blockIdx.x * 64, c_hi = min(c_lo + 64, (g + 1) * Cg);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * B; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const float ga = scal[0], gm = scal[1];
    for (int c = c_lo + warp; c < c_hi; c += kWarpsPerBlock) {
        float gam = gamma[c], sa = 0.f, sb = 0.f;
        for (int b = lane; b < B; b += 32) {
            float4 r = rowsums[(long long)c * B + b];
            float A = ga * r.x + gm * r.z, Bx = ga * r.y + gm * r.w;
            sa += A;
            sb += Bx;
            atomicAdd(&sacc[2 * b], gam * A);
            atomicAdd(&sacc[2 * b + 1], gam * Bx);
        }
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        if (lane == 0) {
            dbeta[c] = sa;
            dgamma[c] = sb;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * B; i += blockDim.x)
        atomicAdd(&S[(size_t)((i >> 1) * G + g) * 2 + (i & 1)], (double)sacc[i]);
}

// backward, step 2: dy = rstd * (gamma * g - m1 - xn * m2), g = (ga loss'(d) + gm 2d)(1 - xh^2)
template <typename OT, bool VEC, bool MSE>
__global__ void __launch_bounds__(kThreads)
recon_bwd_apply_kernel(const float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ x, const float* __restrict__ scal,
                       const double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias, int N, int B, int T,
                       int Tp, int G, int loss_kind, double inv_n) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)N * B;
    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    const float ga = scal[0], gm = scal[1];
    const float g2 = 2.f * (ga + gm);      // MSE: ga * 2d + gm * 2d
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);
        int g = n / (N / G);
        GnStat st = gn_stat(stats, b, g, G, inv_n);
        float gam = gamma[n];
        float a = gam * st.rstd, sh = beta[n] - st.mean * a;
        float m1 = (float)(S[(size_t)(b * G + g) * 2] * inv_n);
        float m2 = (float)(S[(size_t)(b * G + g) * 2 + 1] * inv_n);
        const float* yrow = y + row * Tp;
        long long xo = ((long long)b * N + n) * T;
        float db = 0.f;
        for (int seg = lane; seg < Tp / 8; seg += 32) {
            F8 o;
            if (seg * 8 < T) {
                F8 yv = load8(yrow + seg * 8);
                F8 xv = load8_ext<VEC>(x + xo, seg * 8, T);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float h = tanh_fast(yv.v[i] * a + sh);
                    float d = h - xv.v[i];
                    float gg = (MSE ? g2 * d : ga * loss_grad(loss_kind, d) + gm * 2.f * d) * (1.f - h * h);
                    float xn = (yv.v[i] - st.mean) * st.rstd;
                    float v = (seg * 8 + i < T) ? st.rstd * (gam * gg - m1 - xn * m2) : 0.f;
                    o.v[i] = v;
                    db += v;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
            }
            store8(dy + row * Tp + seg * 8, o);
        }
        db = warp_sum(db);
        if (lane == 0) atomicAdd(&dbias[n], db);
    }
}

static int rows_grid(long long rows) { return (int)cdiv(rows, kWarpsPerBlock); }

template <typename OT>
static int launch_gn_bwd(BwdArgs p, bool loss, int res_is_f32, OT* dy, int planes, long long pstride, float* dgamma,
                         float* dbeta, float* dbias, float* dres, int dres_accumulate, double* ws, cudaStream_t st) {
    size_t sm = planes > 1 ? sizeof(float) * kWarpsPerBlock * (p.Tp + 8) : 0;
    long long rows = (long long)p.C * p.B;
    bool has_gn = p.stats != nullptr;
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * p.C, st);
    if (has_gn) {
        cudaMemsetAsync(dgamma, 0, sizeof(float) * p.C, st);
        cudaMemsetAsync(dbeta, 0, sizeof(float) * p.C, st);
        cudaMemsetAsync(ws, 0, sizeof(double) * 2 * p.B * p.G, st);
    }
    int grid = rows_grid(rows);
#define SG_LAUNCH_BWD(RT, LOSS)                                                                                  \
    do {                                                                                                         \
        if (has_gn) gn_bwd_reduce_kernel<RT, LOSS><<<grid, kThreads, 0, st>>>(p, dgamma, dbeta, ws);             \
        gn_bwd_apply_kernel<OT, RT, LOSS><<<grid, kThreads, sm, st>>>(p, ws, dy, planes, pstride, dbias, dres,   \
                                                                      dres_accumulate);                          \
    } while (0)
    if (loss) {
        SG_LAUNCH_BWD(float, true);
    } else if (p.res == nullptr || res_is_f32) {
        SG_LAUNCH_BWD(float, false);
    } else {
        SG_LAUNCH_BWD(OT, false);
    }
#undef SG_LAUNCH_BWD
    return check_launch("gn_act_bwd");
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_pack_input(const float* x, void* out, int B, int N, int T, int Tp, int dtype, void* stream) {
    SG_REQUIRE(Tp % 8 == 0 && Tp >= T, "pack_input: bad Tp=%d for T=%d", Tp, T);
    long long rows = (long long)N * B;
    if (dtype == SG_BF16)
        pack_input_kernel<__nv_bfloat16><<<rows_grid(rows), kThreads, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)out, B, N, T, Tp);
    else
        pack_input_kernel<float><<<rows_grid(rows), kThreads, 0, as_stream(stream)>>>(x, (float*)out, B, N, T, Tp);
    return check_launch("pack_input");
}

int sg_unpack_f32(const float* in, float* out, int B, int C, int T, int Tp, void* stream) {
    unpack_f32_kernel<<<rows_grid((long long)C * B), kThreads, 0, as_stream(stream)>>>(in, out, B, C, T, Tp);
    return check_launch("unpack_f32");
}

int sg_axpy_f32(float* dst, const float* src, float alpha, long long n, int accumulate, void* stream) {
    if (n <= 0) return 0;
    int grid = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
    axpy_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(dst, src, alpha, n, accumulate);
    return check_launch("axpy_f32");
}

int sg_cast_f32(const float* in, void* out, long long n, int dtype, void* stream) {
    if (n <= 0) return 0;
    int grid = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
    if (dtype == SG_BF16)
        cast_f32_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(in, (__nv_bfloat16*)out, n);
    else
        cast_f32_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(in, (float*)out, n);
    return check_launch("cast_f32");
}

int sg_scale_f64_to_f32(const double* in, float* out, double scale, int n, void* stream) {
    scale_f64_to_f32_kernel<<<(n + 63) / 64, 64, 0, as_stream(stream)>>>(in, out, scale, n);
    return check_launch("scale_f64_to_f32");
}

int sg_gn_stats(const float* y, double* stats, int C, int B, int T, int Tp, int G, void* stream) {
    SG_REQUIRE(G > 0 && C % G == 0, "gn_stats: C=%d not divisible by G=%d", C, G);
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * G, st);
    int Cg = C / G;
    int rpb = Cg < 64 ? Cg : 64;
    dim3 grid((unsigned)cdiv(Cg, rpb), G, B);
    gn_stats_kernel<<<grid, kThreads, 0, st>>>(y, stats, C, B, T, Tp, G, rpb);
    return check_launch("gn_stats");
}

int sg_gn_act_fwd(const float* y, const double* stats, const float* gamma, const float* beta, const void* res,
                  int res_is_f32, float res_scale, int act, int post_gelu, void* out_op, int planes,
                  long long plane_stride, float* out_f32, int C, int B, int T, int Tp, int G, int dtype, void* stream) {
    SG_REQUIRE(Tp % 8 == 0, "gn_act_fwd: Tp %% 8 != 0");
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "gn_act_fwd: planes must be 1, 3 or 5");
    size_t sm = (planes > 1 && out_op) ? sizeof(float) * kWarpsPerBlock * (Tp + 8) : 0;
    SG_REQUIRE(stats == nullptr || (G > 0 && C % G == 0), "gn_act_fwd: bad groups");
    cudaStream_t st = as_stream(stream);
    long long rows = (long long)C * B;
    double inv_n = stats ? 1.0 / ((double)(C / G) * T) : 0.0;
    int grid = rows_grid(rows);
    if (G <= 0) G = 1;
#define SG_FWD(OT, RT)                                                                                            \
    gn_act_fwd_kernel<OT, RT><<<grid, kThreads, sm, st>>>(y, stats, gamma, beta, (const RT*)res, res_scale, act,  \
                                                          post_gelu, (OT*)out_op, planes, plane_stride, out_f32,  \
                                                          C, B, T, Tp, G, inv_n)
    if (dtype == SG_BF16) {
        if (res == nullptr || res_is_f32) SG_FWD(__nv_bfloat16, float);
        else SG_FWD(__nv_bfloat16, __nv_bfloat16);
    } else {
        SG_FWD(float, float);
    }
#undef SG_FWD
    return check_launch("gn_act_fwd");
}

int sg_gn_act_bwd(const float* y, const double* stats, const float* gamma, const float* beta, const void* res,
                  int res_is_f32, float res_scale, int act, int post_gelu, const float* dout, void* dy, int planes,
                  long long plane_stride, float* dgamma, float* dbeta, float* dbias, float* dres, int dres_accumulate,
                  double* ws, int C, int B, int T, int Tp, int G, int dtype, void* stream) {
    SG_REQUIRE(Tp % 8 == 0, "gn_act_bwd: Tp %% 8 != 0");
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "gn_act_bwd: planes must be 1, 3 or 5");
    SG_REQUIRE(stats == nullptr || (G > 0 && C % G == 0 && dgamma && dbeta && ws), "gn_act_bwd: bad GN arguments");
    if (G <= 0) G = 1;
    BwdArgs p{};
    p.y = y; p.stats = stats; p.gamma = gamma; p.beta = beta; p.res = res; p.res_scale = res_scale;
    p.act = act; p.post_gelu = post_gelu; p.dout = dout; p.x = nullptr; p.ext = nullptr; p.ga = 0; p.gm = 0;
    p.loss_kind = 0; p.C = C; p.B = B; p.T = T; p.Tp = Tp; p.G = G;
    p.inv_n = stats ? 1.0 / ((double)(C / G) * T) : 0.0;
    if (dtype == SG_BF16)
        return launch_gn_bwd<__nv_bfloat16>(p, false, res_is_f32, (__nv_bfloat16*)dy, planes, plane_stride, dgamma, dbeta,
                                            dbias, dres, dres_accumulate, ws, as_stream(stream));
    return launch_gn_bwd<float>(p, false, 1, (float*)dy, planes, plane_stride, dgamma, dbeta, dbias, dres,
                                dres_accumulate, ws, as_stream(stream));
}

static int persistent_grid(long long rows) {
    long long blocks = cdiv(rows, kWarpsPerBlock);
    long long cap = 148LL * 16;
    return (int)(blocks < cap ? blocks : cap);
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int sg_recon_fwd(const float* y, const double* stats, const float* gamma, const float* beta, const float* x,
                 float* x_hat, double* loss_sums, float* rowsums, int N, int B, int T, int Tp, int G, int loss_kind,
                 void* stream) {
    SG_REQUIRE(G > 0 && N % G == 0, "recon_fwd: N=%d not divisible by G=%d", N, G);
    SG_REQUIRE(rowsums == nullptr || (x != nullptr && aligned16(rowsums)), "recon_fwd: rowsums needs x and 16-byte alignment");
    cudaStream_t st = as_stream(stream);
    if (x != nullptr) cudaMemsetAsync(loss_sums, 0, sizeof(double) * 2, st);
    double inv_n = 1.0 / ((double)(N / G) * T);
    bool vec = (T % 4 == 0) && aligned16(x) && aligned16(x_hat);
    int grid = persistent_grid((long long)N * B);
    const bool mse = loss_kind == SG_LOSS_MSE;
#define SG_RFWD(VEC, MSE)                                                                                              \
    recon_fwd_kernel<VEC, MSE><<<grid, kThreads, 0, st>>>(y, stats, gamma, beta, x, x_hat, loss_sums, (float4*)rowsums, N, \
                                                          B, T, Tp, G, loss_kind, inv_n)
    if (vec) { if (mse) SG_RFWD(true, true); else SG_RFWD(true, false); }
    else     { if (mse) SG_RFWD(false, true); else SG_RFWD(false, false); }
#undef SG_RFWD
    return check_launch("recon_fwd");
}

// g_loss / g_mse are device scalars produced by autograd (may be NULL); fold them with inv_numel on device.
__global__ void recon_scalars_kernel(const float* g_loss, const float* g_mse, float inv_numel, float* out2) {
    out2[0] = g_loss ? g_loss[0] * inv_numel : 0.f;
    out2[1] = g_mse ? g_mse[0] * inv_numel : 0.f;
}

}  // extern "C"

namespace sg {
// The loss-derived backward needs ga/gm as kernel *values*; they live on the device (autograd
// scalars), so the two recon backward kernels read them through this small indirection.
template <typename OT, bool REDUCE>
__global__ void __launch_bounds__(kThreads)
recon_bwd_kernel(BwdArgs p, const float* __restrict__ scal, float* __restrict__ dgamma, float* __restrict__ dbeta,
                 double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias) {
    p.ga = scal[0];
    p.gm = scal[1];
    long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= (long long)p.C * p.B) return;
    int lane = threadIdx.x & 31;
    int c = (int)(row / p.B), b = (int)(row % p.B);
    int g = c / (p.C / p.G);
    GnStat st = gn_stat(p.stats, b, g, p.G, p.inv_n);
    float gm = p.gamma[c];
    float a = gm * st.rstd, sh = p.beta[c] - st.mean * a;
    if (REDUCE) {
        float A = 0.f, Bx = 0.f;
        for (int seg = lane; seg < p.Tp / 8; seg += 32) {
            F8 dyh, xhat, dpre;
            bwd_segment<float, true>(p, row, c, b, seg, a, sh, st.mean, st.rstd, dyh, xhat, dpre);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                A += dyh.v[i];
                Bx += dyh.v[i] * xhat.v[i];
            }
        }
        A = warp_sum(A);
        Bx = warp_sum(Bx);
        if (lane == 0) {
            atomicAdd(&dgamma[c], Bx);
            atomicAdd(&dbeta[c], A);
            atomicAdd(&S[(size_t)(b * p.G + g) * 2], (double)(gm * A));
            atomicAdd(&S[(size_t)(b * p.G + g) * 2 + 1], (double)(gm * Bx));
        }
    } else {
        float m1 = (float)(S[(size_t)(b * p.G + g) * 2] * p.inv_n);
        float m2 = (float)(S[(size_t)(b * p.G + g) * 2 + 1] * p.inv_n);
        float db = 0.f;
        for (int seg = lane; seg < p.Tp / 8; seg += 32) {
            F8 dyh, xhat, dpre, o;
            bwd_segment<float, true>(p, row, c, b, seg, a, sh, st.mean, st.rstd, dyh, xhat, dpre);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = (seg * 8 + i < p.T) ? st.rstd * (gm * dyh.v[i] - m1 - xhat.v[i] * m2) : 0.f;
                o.v[i] = v;
                db += v;
            }
            store8(dy + row * p.Tp + seg * 8, o);
        }
        db = warp_sum(db);
        if (lane == 0) atomicAdd(&dbias[c], db);
    }
}
}  // namespace sg

extern "C" int sg_recon_bwd(const float* y, const double* stats, const float* gamma, const float* beta,
                            const float* x, const float* g_loss, const float* g_mse, float inv_numel,
                            const float* dxhat_ext, const float* rowsums, void* dy, float* dgamma, float* dbeta,
                            float* dbias, double* ws, int N, int B, int T, int Tp, int G, int loss_kind, int dtype,
                            void* stream) {
    SG_REQUIRE(G > 0 && N % G == 0 && Tp % 8 == 0, "recon_bwd: bad shape");
    SG_REQUIRE(x != nullptr || (g_loss == nullptr && g_mse == nullptr), "recon_bwd: loss gradient without x");
    cudaStream_t st = as_stream(stream);
    double inv_n = 1.0 / ((double)(N / G) * T);
    // workspace: 2*B*G doubles for S followed by 2 floats for the folded scalars
    double* S = ws;
    float* scal = reinterpret_cast<float*>(ws + 2 * (size_t)B * G);
    cudaMemsetAsync(S, 0, sizeof(double) * 2 * B * G, st);
    cudaMemsetAsync(dbias, 0, sizeof(float) * N, st);
    recon_scalars_kernel<<<1, 1, 0, st>>>(g_loss, g_mse, inv_numel, scal);
    if (rowsums != nullptr && dxhat_ext == nullptr && x != nullptr) {
        // one pass over y / x: the reductions of the GroupNorm backward were taken by the forward
        SG_REQUIRE((size_t)B * 2 * sizeof(float) <= 48 * 1024, "recon_bwd: batch too large for the combine kernel");
        dim3 gc((unsigned)cdiv(N / G, 64), G);
        recon_bwd_combine_kernel<<<gc, kThreads, sizeof(float) * 2 * B, st>>>((const float4*)rowsums, scal, gamma, dgamma,
                                                                               dbeta, S, N, B, G);
        int grid = persistent_grid((long long)N * B);
        bool vec = (T % 4 == 0) && aligned16(x);
        const bool mse = loss_kind == SG_LOSS_MSE;
#define SG_APPLY(OT, VEC, MSE)                                                                                       \
    recon_bwd_apply_kernel<OT, VEC, MSE><<<grid, kThreads, 0, st>>>(y, stats, gamma, beta, x, scal, S, (OT*)dy, dbias, N, \
                                                                     B, T, Tp, G, loss_kind, inv_n)
#define SG_APPLY2(OT, VEC) do { if (mse) SG_APPLY(OT, VEC, true); else SG_APPLY(OT, VEC, false); } while (0)
        if (dtype == SG_BF16) { if (vec) SG_APPLY2(__nv_bfloat16, true); else SG_APPLY2(__nv_bfloat16, false); }
        else                  { if (vec) SG_APPLY2(float, true); else SG_APPLY2(float, false); }
#undef SG_APPLY2
#undef SG_APPLY
        return check_launch("recon_bwd");
    }
    BwdArgs p{};
    p.y = y; p.stats = stats; p.gamma = gamma; p.beta = beta; p.res = nullptr; p.res_scale = 1.f;
    p.act = SG_ACT_TANH; p.post_gelu = 0; p.dout = nullptr; p.x = x; p.ext = dxhat_ext; p.loss_kind = loss_kind;
    p.C = N; p.B = B; p.T = T; p.Tp = Tp; p.G = G;
    p.inv_n = inv_n;
    cudaMemsetAsync(dgamma, 0, sizeof(float) * N, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * N, st);
    int grid = rows_grid((long long)N * B);
    if (dtype == SG_BF16) {
        recon_bwd_kernel<__nv_bfloat16, true><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (__nv_bfloat16*)dy, dbias);
        recon_bwd_kernel<__nv_bfloat16, false><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (__nv_bfloat16*)dy, dbias);
    } else {
        recon_bwd_kernel<float, true><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (float*)dy, dbias);
        recon_bwd_kernel<float, false><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (float*)dy, dbias);
    }
    return check_launch("recon_bwd");
}
