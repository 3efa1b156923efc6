// Static fields (preset Dim2 = 1 -> num_time = 1, modules/utils.py:306: SimulGen-VAE.py:279-283 feeds [P, N, 1] fields
// through the same Conv1d model) -
// compact forms of the two N-channel layers.
//
// The engine's CR layout pads every row to 8 elements, so at T = 1 the activations of the first encoder conv (N -> C,
// encoder.py:34) and of the reconstruction conv (C -> N, decoder.py:117) carry 1 valid column in 8: the tensor-core GEMMs
// multiply 7 zero columns per sample and the head kernels stream 8x the bytes.  For T = 1 a k = 1 conv is a plain matrix
// product over the batch, out[co][b] = sum_n W[co][n] x[n][b], so these two layers run on COMPACT operands [C][B]
// (B % 8 == 0): to the GEMM entry points a compact tensor is simply an activation with B / 8 "samples" of 8 valid columns.
//   sg_pack_static            x fp32 [B][N] -> xc [N][B] operand format (+ optional fp32 transpose, the loss target)
//   sg_rows_compact16         padded [R][8] -> [R]   (column 0)       16-bit: small C x B tensors next to the big layers
//   sg_rows_expand_f32        [R] -> padded [R][8] (zeros in columns 1..7), fp32, optionally accumulating onto column 0
//   sg_static_stats           GroupNorm(G, N) statistics per (sample, group) of y [N][B]
//   sg_static_recon_fwd       Tanh(GroupNorm(y)) vs x: loss sums + the reductions of the GroupNorm backward
//                             (per channel and per (sample, group)): NO per-row side buffer; optionally x_hat, transposed
//   sg_static_recon_bwd       dy [N][B] (operand format), dgamma, dbeta, dbias
// Thread mapping of the head kernels: a thread owns one OCTET of samples (8 consecutive b: one 16-byte load of y) and walks
// channels of ONE group, so the (mean, rstd) and the per-(sample, group) accumulators of its 8 samples stay in registers;
// the B / 8 threads of a channel read B contiguous elements.
#include "common.cuh"

namespace sg {

constexpr int kStThreads = 256;
typedef __nv_bfloat16 h16;

static bool st_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- layout ------------------------------------------------------------------------------------------------------
// tile: 64 nodes x 64 samples.  Reads walk a sample's 64 contiguous nodes (256 bytes per warp and row: 128-byte pieces ran
// at 2.7 TB/s, profiles/r2_ncu_config4_static_summary.txt), writes walk a node's 64 contiguous samples.
template <typename OT>
__global__ void __launch_bounds__(kStThreads)
pack_static_kernel(const float* __restrict__ x, OT* __restrict__ xc, float* __restrict__ xt, int B, int N) {
    __shared__ float tile[64][65];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * 64, b0 = blockIdx.y * 64;
    const bool pairs = (N & 1) == 0;                         // 8-byte aligned pairs of nodes
    for (int i = warp; i < 64; i += kStThreads / 32) {
        const int b = b0 + i, n = n0 + 2 * lane;
        float v0 = 0.f, v1 = 0.f;
        if (b < B) {
            const float* src = x + (size_t)b * N + n;
            if (pairs && n + 1 < N) {
                const float2 v = __ldcs(reinterpret_cast<const float2*>(src));
                v0 = v.x;
                v1 = v.y;
            } else {
                if (n < N) v0 = __ldcs(src);
                if (n + 1 < N) v1 = __ldcs(src + 1);
            }
        }
        tile[i][2 * lane] = v0;
        tile[i][2 * lane + 1] = v1;
    }
    __syncthreads();
    for (int j = warp; j < 64; j += kStThreads / 32) {
        const int n = n0 + j, b = b0 + 2 * lane;             // B % 8 == 0: pairs never straddle the edge
        if (n < N && b < B) {
            const float v0 = tile[2 * lane][j], v1 = tile[2 * lane + 1][j];
            if (sizeof(OT) == 2) *reinterpret_cast<uint32_t*>(xc + (size_t)n * B + b) = f2_to_op16x2(v0, v1);
            else *reinterpret_cast<float2*>(xc + (size_t)n * B + b) = make_float2(v0, v1);
            if (xt != nullptr) *reinterpret_cast<float2*>(xt + (size_t)n * B + b) = make_float2(v0, v1);
        }
    }
}

__global__ void rows_compact16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, long long R) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) out[r] = in[r * 8];
}

__global__ void rows_expand_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long R, int accumulate) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float* o = out + r * 8;
    if (accumulate) {
        o[0] += in[r];
    } else {
        *reinterpret_cast<float4*>(o) = make_float4(in[r], 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(o + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- head kernels: common mapping -----------------------------------------------------------------------------------
struct StMap {
    int octet, nl, rpi, c_lo, c_hi, g;
    bool active, whole_warps;
};
// block (x, g): channels [c_lo, c_hi) of group g; thread: octet of samples, channel phase nl of rpi channels per iteration
__device__ __forceinline__ StMap st_map(int N, int B, int G) {
    StMap m;
    const int OB = B >> 3;
    m.rpi = kStThreads / OB;
    m.octet = threadIdx.x % OB;
    m.nl = threadIdx.x / OB;
    m.active = m.nl < m.rpi;
    m.whole_warps = (OB & 31) == 0;                          // a warp lies inside one channel: warp_sum per channel
    m.g = blockIdx.y;
    const int Cg = N / G;
    const int per = (Cg + gridDim.x - 1) / gridDim.x;
    m.c_lo = m.g * Cg + blockIdx.x * per;
    m.c_hi = min(m.c_lo + per, (m.g + 1) * Cg);
    return m;
}
// sum of v over the threads of one channel, added to dst (one atomic per warp when warps do not straddle channels)
__device__ __forceinline__ void st_channel_add(float* dst, float v, const StMap& m) {
    if (m.whole_warps) {
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
    } else {
        atomicAdd(dst, v);
    }
}
// acc[NV] of every thread summed over the threads with the same octet; the nl == 0 thread of each octet gets the total
template <int NV>
__device__ __forceinline__ void st_octet_reduce(float (&acc)[NV], float* sh, const StMap& m, int OB) {
    __syncthreads();
    if (m.active && m.nl > 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) sh[((m.nl - 1) * OB + m.octet) * NV + k] = acc[k];
    }
    __syncthreads();
    if (m.active && m.nl == 0) {
        for (int r = 1; r < m.rpi; ++r) {
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] += sh[((r - 1) * OB + m.octet) * NV + k];
        }
    }
}

template <typename YT>
__global__ void __launch_bounds__(kStThreads)
static_stats_kernel(const YT* __restrict__ y, double* __restrict__ sums, int N, int B, int G) {
    extern __shared__ float st_sh[];
    const StMap m = st_map(N, B, G);
    const int OB = B >> 3;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    if (m.active) {
#pragma unroll 4
        for (int c = m.c_lo + m.nl; c < m.c_hi; c += m.rpi) {
            const F8 v = load8(y + (size_t)c * B + m.octet * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i] += v.v[i];
                acc[8 + i] = fmaf(v.v[i], v.v[i], acc[8 + i]);
            }
        }
    }
    st_octet_reduce<16>(acc, st_sh, m, OB);
    if (m.active && m.nl == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const size_t o = (size_t)((m.octet * 8 + i) * G + m.g) * 2;
            atomicAdd(&sums[o], (double)acc[i]);
            atomicAdd(&sums[o + 1], (double)acc[8 + i]);
        }
    }
}

__global__ void static_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mr, int n, double inv_n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mean = sums[2 * i] * inv_n;
    double var = sums[2 * i + 1] * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    mr[2 * i] = (float)mean;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
}

// forward: loss sums, chan4[n] = sum_b (aL, bL, aM, bM), s4[b][g] = sum_{n in g} gamma_n (aL, bL, aM, bM) with
//   aM = 2 d (1 - h^2), bM = aM * xhat  (the MSE gradient terms), aL / bL the same with loss'(d) (== the MSE ones for MSE)
template <typename YT, typename XT, bool MSE>
__global__ void __launch_bounds__(kStThreads)           // 80 registers, 3 blocks per SM; forcing 4 spills and is slower (0.86 vs 0.83 ms)
static_recon_fwd_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const XT* __restrict__ x, float* __restrict__ xh,
                        double* __restrict__ loss_sums, float* __restrict__ chan4, float* __restrict__ s4, int N, int B,
                        int G, int loss_kind) {
    extern __shared__ float st_sh[];
    __shared__ double shm[2][32];
    constexpr int NS = MSE ? 2 : 4;
    const StMap m = st_map(N, B, G);
    const int OB = B >> 3;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    float rstd[8], mean[8], acc[8 * NS];
#pragma unroll
    for (int k = 0; k < 8 * NS; ++k) acc[k] = 0.f;
    float l0 = 0.f, l1 = 0.f;
    if (m.active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 st = __ldg(mr2 + (m.octet * 8 + i) * G + m.g);
            mean[i] = st.x;
            rstd[i] = st.y;
        }
#pragma unroll 2
        for (int c = m.c_lo + m.nl; c < m.c_hi; c += m.rpi) {
            const size_t off = (size_t)c * B + m.octet * 8;
            const F8 yv = load8(y + off), xv = load8(x + off);
            const float gam = __ldg(gamma + c), bet = __ldg(beta + c);
            float cs[4] = {0.f, 0.f, 0.f, 0.f};
            F8 hv;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = gam * rstd[i];
                const float h = tanh_fast(fmaf(yv.v[i], a, bet - mean[i] * a));
                hv.v[i] = h;
                const float d = h - xv.v[i];
                l1 = fmaf(d, d, l1);
                if (!MSE) l0 += loss_term(loss_kind, d);
                const float om = fmaf(-h, h, 1.f);
                const float xn = (yv.v[i] - mean[i]) * rstd[i];
                const float am = 2.f * d * om, bm = am * xn;
                cs[2] += am;
                cs[3] += bm;
                acc[i * NS + NS - 2] = fmaf(gam, am, acc[i * NS + NS - 2]);
                acc[i * NS + NS - 1] = fmaf(gam, bm, acc[i * NS + NS - 1]);
                if (!MSE) {
                    const float al = loss_grad(loss_kind, d) * om, bl = al * xn;
                    cs[0] += al;
                    cs[1] += bl;
                    acc[i * NS] = fmaf(gam, al, acc[i * NS]);
                    acc[i * NS + 1] = fmaf(gam, bl, acc[i * NS + 1]);
                }
            }
            if (MSE) { cs[0] = cs[2]; cs[1] = cs[3]; }
            if (xh != nullptr) store8(xh + off, hv);         // x_hat, transposed: [N][B]
#pragma unroll
            for (int k = 0; k < 4; ++k) st_channel_add(chan4 + (size_t)c * 4 + k, cs[k], m);
        }
    }
    st_octet_reduce<8 * NS>(acc, st_sh, m, OB);
    if (m.active && m.nl == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float* o = s4 + (size_t)((m.octet * 8 + i) * G + m.g) * 4;
            atomicAdd(o + 2, acc[i * NS + NS - 2]);
            atomicAdd(o + 3, acc[i * NS + NS - 1]);
            atomicAdd(o + 0, acc[i * NS]);                   // MSE: NS == 2, i.e. the same two values
            atomicAdd(o + 1, acc[i * NS + 1]);
        }
    }
    const double t0 = block_sum((double)(MSE ? l1 : l0), shm[0]);
    const double t1 = block_sum((double)l1, shm[1]);
    if (threadIdx.x == 0) {
        atomicAdd(&loss_sums[0], t0);
        atomicAdd(&loss_sums[1], t1);
    }
}

// upstream scalars folded with 1 / numel; dgamma / dbeta from chan4; S[b][g] = (m1, m2) * Cg from s4
__global__ void static_recon_combine_kernel(const float* __restrict__ g_loss, const float* __restrict__ g_mse, float inv_numel,
                                            const float* __restrict__ chan4, const float* __restrict__ s4,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ S2,
                                            float* __restrict__ scal, int N, int BG) {
    const float ga = g_loss ? g_loss[0] * inv_numel : 0.f, gm = g_mse ? g_mse[0] * inv_numel : 0.f;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { scal[0] = ga; scal[1] = gm; }
    if (i < N) {
        const float4 c = *reinterpret_cast<const float4*>(chan4 + (size_t)i * 4);
        dbeta[i] = ga * c.x + gm * c.z;
        dgamma[i] = ga * c.y + gm * c.w;
    }
    if (i < BG) {
        const float4 s = *reinterpret_cast<const float4*>(s4 + (size_t)i * 4);
        S2[2 * i] = ga * s.x + gm * s.z;
        S2[2 * i + 1] = ga * s.y + gm * s.w;
    }
}

template <typename YT, typename XT, bool MSE>
__global__ void __launch_bounds__(kStThreads, 4)
static_recon_bwd_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const XT* __restrict__ x, const float* __restrict__ scal,
                        const float* __restrict__ S2, h16* __restrict__ dy, float* __restrict__ dbias, int N, int B, int G,
                        int loss_kind, float inv_n) {
    const StMap m = st_map(N, B, G);
    if (!m.active) return;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    const float ga = scal[0], gm = scal[1];
    const float g2 = 2.f * (ga + gm);
    float rstd[8], mean[8], c2[8], c3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int bg = (m.octet * 8 + i) * G + m.g;
        const float2 st = __ldg(mr2 + bg);
        mean[i] = st.x;
        rstd[i] = st.y;
        const float m1 = S2[2 * bg] * inv_n, m2 = S2[2 * bg + 1] * inv_n;
        c2[i] = -st.y * st.y * m2;
        c3[i] = st.y * (st.x * st.y * m2 - m1);
    }
#pragma unroll 2
    for (int c = m.c_lo + m.nl; c < m.c_hi; c += m.rpi) {
        const size_t off = (size_t)c * B + m.octet * 8;
        const F8 yv = load8(y + off), xv = load8(x + off);
        const float gam = __ldg(gamma + c), bet = __ldg(beta + c);
        F8 o;
        float db = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = gam * rstd[i];
            const float h = tanh_fast(fmaf(yv.v[i], a, bet - mean[i] * a));
            const float d = h - xv.v[i];
            const float gg = (MSE ? g2 * d : ga * loss_grad(loss_kind, d) + gm * 2.f * d) * fmaf(-h, h, 1.f);
            o.v[i] = fmaf(a, gg, fmaf(c2[i], yv.v[i], c3[i]));
            db += o.v[i];
        }
        store8(dy + off, o);
        st_channel_add(dbias + c, db, m);
    }
}

// blocks split the channels of a group evenly: launch what is resident (one wave), per kernel instantiation
template <typename KernelT>
static dim3 st_grid(KernelT kernel, size_t smem, int N, int G) {
    const int Cg = N / G;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kStThreads, smem) != cudaSuccess || nb < 1) nb = 2;
    int per_group = 148 * nb / G;
    if (per_group > Cg) per_group = Cg;
    if (per_group < 1) per_group = 1;
    return dim3((unsigned)per_group, (unsigned)G);
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_pack_static(const float* x, void* xc, float* xt, int B, int N, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_REQUIRE(B > 0 && B % 8 == 0 && N > 0, "pack_static: the batch must be a multiple of 8 (B=%d)", B);
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(B, 64));
    cudaStream_t st = as_stream(stream);
    if (is_op16(dtype)) pack_static_kernel<h16><<<grid, kStThreads, 0, st>>>(x, (h16*)xc, xt, B, N);
    else pack_static_kernel<float><<<grid, kStThreads, 0, st>>>(x, (float*)xc, xt, B, N);
    return check_launch("pack_static");
}

int sg_rows_compact16(const void* in, void* out, long long R, void* stream) {
    rows_compact16_kernel<<<(unsigned)cdiv(R, 256), 256, 0, as_stream(stream)>>>((const uint16_t*)in, (uint16_t*)out, R);
    return check_launch("rows_compact16");
}

int sg_rows_expand_f32(const float* in, float* out, long long R, int accumulate, void* stream) {
    SG_REQUIRE(st_aligned16(out), "rows_expand_f32: unaligned output");
    rows_expand_f32_kernel<<<(unsigned)cdiv(R, 256), 256, 0, as_stream(stream)>>>(in, out, R, accumulate);
    return check_launch("rows_expand_f32");
}

#define SG_STATIC_SHAPE(name)                                                                                          \
    SG_REQUIRE(G > 0 && N % G == 0 && B % 8 == 0 && B >= 8 && B <= 8 * kStThreads && st_aligned16(y),                   \
               name ": needs N %% G == 0 and a batch that is a multiple of 8 and <= %d (N=%d G=%d B=%d)", 8 * kStThreads, N, G, B)

int sg_static_stats(const void* y, int y_dtype, double* ws, float* mr, int N, int B, int G, void* stream) {
    SG_CHECK_OP16(y_dtype);
    SG_STATIC_SHAPE("static_stats");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
    const size_t sm = sizeof(float) * kStThreads * 16;
    if (is_op16(y_dtype)) static_stats_kernel<h16><<<st_grid(static_stats_kernel<h16>, sm, N, G), kStThreads, sm, st>>>((const h16*)y, ws, N, B, G);
    else static_stats_kernel<float><<<st_grid(static_stats_kernel<float>, sm, N, G), kStThreads, sm, st>>>((const float*)y, ws, N, B, G);
    static_finalize_kernel<<<(unsigned)cdiv(B * G, 256), 256, 0, st>>>(ws, mr, B * G, 1.0 / (double)(N / G));
    return check_launch("static_stats");
}

// ws: 4 * N + 4 * B * G floats (chan4, s4), zeroed here; kept for sg_static_recon_bwd
int sg_static_recon_fwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* x,
                        int x_dtype, float* xhat_t, double* loss_sums, float* ws, int N, int B, int G, int loss_kind,
                        void* stream) {
    SG_CHECK_OP16(y_dtype);
    SG_CHECK_OP16(x_dtype);
    SG_STATIC_SHAPE("static_recon_fwd");
    SG_REQUIRE(is_op16(y_dtype) && st_aligned16(x) && st_aligned16(ws) && st_aligned16(xhat_t),
               "static_recon_fwd: y must be in the 16-bit operand format");
    cudaStream_t st = as_stream(stream);
    float* chan4 = ws;
    float* s4 = ws + (size_t)4 * N;
    cudaMemsetAsync(ws, 0, sizeof(float) * 4 * ((size_t)N + (size_t)B * G), st);
    cudaMemsetAsync(loss_sums, 0, sizeof(double) * 2, st);
    const bool mse = loss_kind == SG_LOSS_MSE;
#define SG_SF(XT, MSE) static_recon_fwd_kernel<h16, XT, MSE><<<st_grid(static_recon_fwd_kernel<h16, XT, MSE>, sizeof(float) * kStThreads * 8 * (MSE ? 2 : 4), N, G), \
                                                              kStThreads, sizeof(float) * kStThreads * 8 * (MSE ? 2 : 4), st>>>( \
        (const h16*)y, mr, gamma, beta, (const XT*)x, xhat_t, loss_sums, chan4, s4, N, B, G, loss_kind)
    if (is_op16(x_dtype)) { if (mse) SG_SF(h16, true); else SG_SF(h16, false); }
    else                  { if (mse) SG_SF(float, true); else SG_SF(float, false); }
#undef SG_SF
    return check_launch("static_recon_fwd");
}

// ws: what sg_static_recon_fwd left, followed by 2 * B * G + 2 floats of scratch
int sg_static_recon_bwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* x,
                        int x_dtype, const float* g_loss, const float* g_mse, float inv_numel, float* ws, void* dy,
                        float* dgamma, float* dbeta, float* dbias, int N, int B, int G, int loss_kind, int dtype, void* stream) {
    SG_CHECK_OP16(y_dtype);
    SG_CHECK_OP16(x_dtype);
    SG_CHECK_OP16(dtype);
    SG_STATIC_SHAPE("static_recon_bwd");
    SG_REQUIRE(is_op16(y_dtype) && is_op16(dtype) && st_aligned16(x) && st_aligned16(dy) && st_aligned16(ws),
               "static_recon_bwd: y and dy must be in the 16-bit operand format");
    cudaStream_t st = as_stream(stream);
    float* chan4 = ws;
    float* s4 = ws + (size_t)4 * N;
    float* S2 = s4 + (size_t)4 * B * G;
    float* scal = S2 + (size_t)2 * B * G;
    const int n = N > B * G ? N : B * G;
    static_recon_combine_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(g_loss, g_mse, inv_numel, chan4, s4, dgamma, dbeta, S2,
                                                                       scal, N, B * G);
    cudaMemsetAsync(dbias, 0, sizeof(float) * N, st);
    const bool mse = loss_kind == SG_LOSS_MSE;
    const float inv_n = (float)(1.0 / (double)(N / G));
#define SG_SB(XT, MSE) static_recon_bwd_kernel<h16, XT, MSE><<<st_grid(static_recon_bwd_kernel<h16, XT, MSE>, 0, N, G), kStThreads, 0, st>>>( \
        (const h16*)y, mr, gamma, beta, (const XT*)x, scal, S2, (h16*)dy, dbias, N, B, G, loss_kind, inv_n)
    if (is_op16(x_dtype)) { if (mse) SG_SB(h16, true); else SG_SB(h16, false); }
    else                  { if (mse) SG_SB(float, true); else SG_SB(float, false); }
#undef SG_SB
    return check_launch("static_recon_bwd");
}

}  // extern "C"
