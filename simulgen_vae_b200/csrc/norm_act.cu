// HBM-streaming kernels: layout packing, GroupNorm statistics / apply / backward, activation,
// residual, and the reconstruction head (tanh + fused loss reductions).
//
// Reference arithmetic replaced (all fp32 ATen kernels there):
//   encoder.py:35-36, common.py:85-102,108-125,133-162 (GroupNorm -> GELU -> x + 0.1 f(x)),
//   decoder.py:32 (GELU after ConvTranspose1d), decoder.py:117-121 (GroupNorm -> Tanh),
//   VAE_network.py:71-77,110-111 (MSE / L1 / SmoothL1 / Huber, mean reduction) and their backward.
//
// Structure shared by every kernel: a row is the Tp contiguous elements of one (channel, sample) pair, of
// which the first T are valid.  Grids are persistent (a few CTAs per SM); each warp strides over rows, each
// lane owns 8-element (16/32-byte) segments, so global accesses are coalesced 128-bit transactions;
// reductions use warp shuffles + one atomic (or plain store) per row.  GroupNorm statistics reach these
// kernels as fp32 (mean, rstd) pairs per (sample, group) - the fp64 finalisation runs once per layer in
// gn_finalize_kernel, not once per row - and the activation is a template parameter, so the inner loops are
// ~15-30 issue slots per element: below the HBM time per element (round 1: they were ALU-bound at 30-60 %
// of the HBM roofline, gpurun_out/stream_bench32.txt).
#include <stdlib.h>

#include "common.cuh"

namespace sg {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

static int rows_grid(long long rows) { return (int)cdiv(rows, kWarpsPerBlock); }
static int persistent_grid(long long rows) {
    long long blocks = cdiv(rows, kWarpsPerBlock);
    long long cap = 148LL * 8;
    return (int)(blocks < cap ? blocks : cap);
}
// Persistent short-row kernels (Tp == 8) split their tiles evenly over the grid, so the grid must be exactly what is
// RESIDENT: with 888 blocks launched and 4-5 per SM resident (registers), the last 148-296 blocks ran alone after the
// first wave and held a sixth to a third of the work (ncu, profiles/r2_ncu_config4_short_summary.txt: 39-43 % of the
// warp slots active on average).  The occupancy query is per kernel instantiation.
template <typename KernelT>
static int resident_grid(KernelT kernel, long long work_blocks, int max_per_sm = 16) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
    if (nb > max_per_sm) nb = max_per_sm;
    const long long cap = 148LL * nb;
    return (int)(work_blocks < cap ? work_blocks : cap);
}
// warp-per-task persistent kernels (GroupNorm / activation): fixed 148 x 8 grid.  SIMULGEN_B200_GN_GRID=1 launches exactly
// the resident grid for layers with many tasks per warp instead (no partial last wave) - measured on a B200 at the
// headline shapes: no consistent gain (5-plane 16-bit backward 0.321 -> 0.347 ms, 1-plane fp32 backward 0.304 -> 0.288 ms,
// step 38.0 / 38.3 vs 38.5 / 38.4 ms; profiles/r2_gn_grid_ab.txt), so it stays off.
template <typename KernelT>
static int task_grid(KernelT kernel, long long tasks, size_t smem) {
    static const int mode = [] { const char* e = getenv("SIMULGEN_B200_GN_GRID"); return e ? atoi(e) : 0; }();
    const int wide = persistent_grid(tasks);
    if (mode == 0) return wide;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, smem) != cudaSuccess || nb < 1) return wide;
    const long long resident = 148LL * nb;
    if (resident < wide && tasks >= 4 * resident * kWarpsPerBlock) return (int)resident;
    return wide;
}
static long long short_tiles(int N, int B) { return cdiv((long long)N, 32) * cdiv((long long)B, 32); }
// short-row GroupNorm kernels: one warp per (<= 8 channels of a group, 32 samples) task
static long long short_task_blocks(int C, int B, int G) {
    const int Cg = C / G;
    return cdiv((long long)G * cdiv(Cg, 8) * cdiv(B, 32), kWarpsPerBlock);
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// layout kernels
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

template <typename OT, bool VEC>
__global__ void __launch_bounds__(kThreads) pack_input_kernel(const float* __restrict__ x, OT* __restrict__ out,
                                                              int B, int N, int T, int Tp) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)N * B, wstride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);                                 // row = n * B + b
        const float* src = x + ((long long)b * N + n) * T;
        OT* dst = out + row * Tp;
        for (int seg = lane; seg < Tp / 8; seg += 32) store8(dst + seg * 8, load8_ext<VEC>(src, seg * 8, T));
    }
}

// the same re-layout for a batch that is already in the 16-bit operand format ([B][N][T], e.g. a host staging buffer
// kept in fp16 so that the host -> device copy moves half the bytes); T % 8 == 0: whole 16-byte segments
__global__ void __launch_bounds__(kThreads) pack_input16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                                int B, int N, int T) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)N * B, wstride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);
        const uint4* src = reinterpret_cast<const uint4*>(x + ((long long)b * N + n) * T);
        uint4* dst = reinterpret_cast<uint4*>(out + row * T);
        for (int seg = lane; seg < T / 8; seg += 32) dst[seg] = __ldg(src + seg);
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
// GroupNorm statistics: sums[b][g] += (sum, sumsq) over the group's rows; then (mean, rstd) in fp32
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const float* __restrict__ y, double* __restrict__ sums,
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
        atomicAdd(&sums[(size_t)(b * G + g) * 2], t0);
        atomicAdd(&sums[(size_t)(b * G + g) * 2 + 1], t1);
    }
}

// mr[b][g] = (mean, 1 / sqrt(biased var + eps)) - nn.GroupNorm semantics, finalised in fp64
__global__ void gn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mr, int n, double inv_n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m = sums[2 * i] * inv_n;
    double var = sums[2 * i + 1] * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    mr[2 * i] = (float)m;
    mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
}

// ---------------------------------------------------------------------------------------------
// activation dispatch (compile time)
// ---------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
    if (ACT == SG_ACT_GELU) return gelu_f(x);
    if (ACT == SG_ACT_TANH) return tanh_fast(x);
    return x;
}
template <int ACT>
__device__ __forceinline__ void act_both_t(float x, float& val, float& dval) {
    if (ACT == SG_ACT_GELU) {
        gelu_both(x, val, dval);
    } else if (ACT == SG_ACT_TANH) {
        val = tanh_fast(x);
        dval = 1.0f - val * val;
    } else {
        val = x;
        dval = 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward apply: pre = res + res_scale * act(a*y + sh); out = POST ? gelu(pre) : pre
// ---------------------------------------------------------------------------------------------
// Work decomposition of the GroupNorm / activation kernels: a warp takes one (channel, chunk of <= kChunkB samples)
// task and walks the chunk's rows - contiguous in memory - several at a time: every load of R rows is issued before the
// first dependent instruction.  What bounds these kernels is bytes in flight per SM, not issue slots: with one 200-element
// row per warp (400 B of 16-bit loads) and ~32 resident warps an SM has 13 KB outstanding against the ~40 KB that HBM
// latency x bandwidth asks for (round 2, gpurun_out/r2_stream_gn_a.txt: the 16-bit y did not speed the one-row kernel up
// at all).  R = 4 rows for 16-bit streams, 2 for fp32 ones (register budget).  Channel constants are loaded once per
// task, there is no per-row index division, and per-channel sums cost one atomic per task.
constexpr int kChunkB = 16;
template <typename A, typename B2>
struct RowsInFlight {
    static constexpr int value = (sizeof(A) == 2 && sizeof(B2) == 2) ? 4 : 2;
};

template <typename OT, typename RT, typename YT, int ACT, bool POST, int PLANES>
__global__ void __launch_bounds__(kThreads, 3)
gn_act_fwd_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const RT* __restrict__ res, float res_scale, OT* __restrict__ out_op,
                  long long pstride, float* __restrict__ out_f32, int C, int B, int T, int Tp, int G) {
    // multi-plane writers hold 12 neighbour values per row for the shifted copies: 2 rows in flight keep them in registers
    // (measured, gpurun_out/r2_stream_gn_{a,c}.txt: 5 planes 0.185 ms with 1 row, 0.211 ms with 4 rows and spills)
    constexpr int R = PLANES > 1 ? 2 : RowsInFlight<YT, YT>::value;
    constexpr int planes = PLANES;
    extern __shared__ float sg_rows[];
    const int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (Tp + 8);
    const bool multi = PLANES > 1 && out_op != nullptr;
    const int nseg_p = Tp >> 3;
    const bool shfl = multi && nseg_p <= 32;
    if (multi && !shfl) srow_clear_halo(srow, Tp, lane);
    const bool has_gn = mr != nullptr;
    const int Cg = has_gn ? C / G : 1;
    const int nchunk = (B + kChunkB - 1) / kChunkB;
    const int tasks = C * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const int c = task / nchunk, ch = task - c * nchunk;
        const int b_lo = ch * kChunkB, b_hi = min(B, b_lo + kChunkB);
        const int g = has_gn ? c / Cg : 0;
        const float gm = has_gn ? __ldg(gamma + c) : 1.f, bt = has_gn ? __ldg(beta + c) : 0.f;
        for (int b0 = b_lo; b0 < b_hi; b0 += R) {
            float a[R], sh[R];
            bool ok[R];
            long long row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < b_hi;
                const int bb = ok[r] ? b0 + r : b0;
                row[r] = ((long long)c * B + bb) * Tp;
                a[r] = 1.f;
                sh[r] = 0.f;
                if (has_gn) {
                    const float2 st = __ldg(mr2 + bb * G + g);
                    a[r] = gm * st.y;
                    sh[r] = bt - st.x * a[r];
                }
            }
            auto finish = [&](int r, int seg, const F8& yv, const F8& rv, F8& o) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float pre = res_scale * act_t<ACT>(fmaf(yv.v[i], a[r], sh[r]));
                    if (res != nullptr) pre += rv.v[i];
                    if (POST) pre = gelu_f(pre);
                    o.v[i] = pre;
                }
                if (seg * 8 + 8 > T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (seg * 8 + i >= T) o.v[i] = 0.f;
                }
            };
            if (shfl || !multi) {
                // rows of <= 256 elements with shifted planes (one segment per lane, planes from registers + shuffles),
                // or single-plane / fp32-only rows of any length
                for (int seg = lane; seg < (shfl ? 32 : nseg_p); seg += 32) {
                    const bool live = seg * 8 < T;
                    typename RawOf<YT>::type yw[R];
                    typename RawOf<RT>::type rw[R];
                    if (live) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {        // every load of every row first, kept unconverted
                            yw[r] = load_raw(y + row[r] + seg * 8);
                            if (res != nullptr) rw[r] = load_raw(res + row[r] + seg * 8);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        F8 o;
                        if (live) {
                            F8 rv;
                            if (res != nullptr) rv = cvt8(rw[r]);
                            finish(r, seg, cvt8(yw[r]), rv, o);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
                        }
                        if (ok[r]) {
                            if (out_f32 != nullptr && seg < nseg_p) store8(out_f32 + row[r] + seg * 8, o);
                            if (PLANES > 1 && shfl) store_planes_shfl<OT, PLANES>(out_op, row[r], pstride, o, T, nseg_p, lane);
                            else if (out_op != nullptr && seg < nseg_p) store8(out_op + row[r] + seg * 8, o);
                        }
                    }
                }
            } else if (PLANES > 1) {
                // long rows with shifted planes: staged through shared memory, one row at a time
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    if (!ok[r]) continue;
                    for (int seg = lane; seg < nseg_p; seg += 32) {
                        F8 o;
                        if (seg * 8 < T) {
                            F8 yv = load8(y + row[r] + seg * 8), rv;
                            if (res != nullptr) rv = load8(res + row[r] + seg * 8);
                            finish(r, seg, yv, rv, o);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
                        if (out_f32 != nullptr) store8(out_f32 + row[r] + seg * 8, o);
                    }
                    __syncwarp();
                    store_row_planes(out_op, row[r], planes, pstride, srow, T, Tp, lane);
                    __syncwarp();
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  The incoming gradient is a tensor (generic layers); the recon head has its own kernels below.
// ---------------------------------------------------------------------------------------------

struct BwdArgs {
    const float* y;
    const float* mr;          // (mean, rstd) per (b, g) or NULL
    const float* gamma;
    const float* beta;
    const void* res;
    float res_scale;
    const float* dout;        // [C][B][Tp]
    int C, B, T, Tp, G;
    double inv_n;
};

// One fully/partially valid 8-element segment, split into a load phase and a compute phase so that the loads of
// several rows can be issued before the first dependent instruction.
struct SegIn {
    F8 yv, dv, rv;
};
template <typename RT, bool POST>
__device__ __forceinline__ void bwd_load(const BwdArgs& p, long long row, int seg, SegIn& in) {
    in.yv = load8(p.y + row * p.Tp + seg * 8);
    in.dv = load8(p.dout + row * p.Tp + seg * 8);
    const RT* res = reinterpret_cast<const RT*>(p.res);
    if (POST && res != nullptr) in.rv = load8(res + row * p.Tp + seg * 8);
}
// dyh = dL/d(gamma*xhat+beta), xhat, dpre (gradient wrt the residual input)
template <int ACT, bool POST>
__device__ __forceinline__ void bwd_compute(const BwdArgs& p, const SegIn& in, int seg, float a, float sh, float mean, float rstd,
                                            F8& dyh, F8& xhat, F8& dpre) {
    const bool has_res = p.res != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float yh = fmaf(in.yv.v[i], a, sh);
        float val, dval;
        act_both_t<ACT>(yh, val, dval);
        float dp = in.dv.v[i];
        if (POST) {
            float pre = p.res_scale * val + (has_res ? in.rv.v[i] : 0.f);
            dp *= gelu_grad_f(pre);
        }
        dyh.v[i] = p.res_scale * dp * dval;
        xhat.v[i] = (in.yv.v[i] - mean) * rstd;
        dpre.v[i] = dp;
    }
    if (seg * 8 + 8 > p.T) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (seg * 8 + i >= p.T) { dyh.v[i] = 0.f; xhat.v[i] = 0.f; dpre.v[i] = 0.f; }
    }
}

constexpr int kRowsInFlight = 2;      // recon head and no-GroupNorm backward kernels (fp32 streams)

// ---- GroupNorm layers: two passes with a 16-bit-friendly hand-off --------------------------------------------------
// pass 1 reads y and the incoming gradient, evaluates the activation derivative ONCE and leaves
//     dz = res_scale * dp * act'(gamma * xhat + beta),   dp = dout (* gelu'(pre) for the trailing GELU)
// in place of dout (same dtype: fp32, or the 16-bit operand format when the producing dgrad GEMM stored 16 bits),
// together with the reductions of the GroupNorm backward (S[b][g] += gamma_c * (sum dz, sum dz * xhat), dgamma, dbeta)
// and the gradient of the residual input (dres (+)= dp).  pass 2 is then three FMAs per element,
//     dy = rstd * (gamma * dz - m1 - xhat * m2) = c1 * dz + c2 * y + c3,
// plus the shifted operand planes and dbias.  Round 1 read y (fp32) and dout (fp32) twice and evaluated the exact-erf
// GELU derivative in both passes (profiles/r1_ncu_source_gn_bwd_summary.txt: 43 instructions per element, issue-bound).
struct Pass1Args {
    const void* y;
    void* dout;               // in: dout, out: dz
    const float* mr;
    const float* gamma;
    const float* beta;
    const void* res;
    float res_scale;
    int C, B, T, Tp, G;
};

template <typename RT, typename YT, typename DT, int ACT, bool POST>
__global__ void __launch_bounds__(kThreads, 3)
gn_bwd_pass1_kernel(Pass1Args p, float* __restrict__ dgamma, float* __restrict__ dbeta, double* __restrict__ S,
                    float* __restrict__ dres, int dres_accumulate) {
    constexpr int R = RowsInFlight<YT, DT>::value;
    const int lane = threadIdx.x & 31;
    const int Cg = p.C / p.G;
    const int nchunk = (p.B + kChunkB - 1) / kChunkB;
    const int tasks = p.C * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(p.mr);
    const YT* y = reinterpret_cast<const YT*>(p.y);
    DT* dout = reinterpret_cast<DT*>(p.dout);
    const RT* res = reinterpret_cast<const RT*>(p.res);
    const bool has_res = res != nullptr;
    const int nseg_p = p.Tp >> 3;
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const int c = task / nchunk, ch = task - c * nchunk;
        const int b_lo = ch * kChunkB, b_hi = min(p.B, b_lo + kChunkB);
        const int g = c / Cg;
        const float gm = __ldg(p.gamma + c), bt = __ldg(p.beta + c);
        float sumA = 0.f, sumB = 0.f;
        for (int b0 = b_lo; b0 < b_hi; b0 += R) {
            float a[R], sh[R], nm[R], rstd[R], A[R], Bx[R];
            bool ok[R];
            long long row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < b_hi;
                const int bb = ok[r] ? b0 + r : b0;
                row[r] = ((long long)c * p.B + bb) * p.Tp;
                const float2 st = __ldg(mr2 + bb * p.G + g);
                rstd[r] = st.y;
                a[r] = gm * st.y;
                sh[r] = bt - st.x * a[r];
                nm[r] = -st.x * st.y;                 // xhat = y * rstd + nm
                A[r] = 0.f;
                Bx[r] = 0.f;
            }
            for (int seg = lane; seg < nseg_p; seg += 32) {
                const bool live = seg * 8 < p.T;
                typename RawOf<YT>::type yw[R];
                typename RawOf<DT>::type dw[R];
                typename RawOf<RT>::type rw[R];
                if (live) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {        // every load of every row first, kept unconverted
                        yw[r] = load_raw(y + row[r] + seg * 8);
                        dw[r] = load_raw(dout + row[r] + seg * 8);
                        if (POST && has_res) rw[r] = load_raw(res + row[r] + seg * 8);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    F8 dz, dp;
                    if (live) {
                        const F8 yv = cvt8(yw[r]), dv = cvt8(dw[r]);
                        F8 rv;
                        if (POST && has_res) rv = cvt8(rw[r]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float val, dval;
                            act_both_t<ACT>(fmaf(yv.v[i], a[r], sh[r]), val, dval);
                            float d = dv.v[i];
                            if (POST) d *= gelu_grad_f(fmaf(p.res_scale, val, has_res ? rv.v[i] : 0.f));
                            dp.v[i] = d;
                            dz.v[i] = p.res_scale * d * dval;
                        }
                        if (seg * 8 + 8 > p.T) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (seg * 8 + i >= p.T) { dz.v[i] = 0.f; dp.v[i] = 0.f; }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            A[r] += dz.v[i];
                            Bx[r] = fmaf(dz.v[i], fmaf(yv.v[i], rstd[r], nm[r]), Bx[r]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { dz.v[i] = 0.f; dp.v[i] = 0.f; }
                    }
                    if (ok[r]) {
                        store8(dout + row[r] + seg * 8, dz);
                        if (dres != nullptr) {
                            float* dr = dres + row[r] + seg * 8;
                            if (dres_accumulate) {
                                F8 old = load8(dr);
#pragma unroll
                                for (int i = 0; i < 8; ++i) dp.v[i] += old.v[i];
                            }
                            store8(dr, dp);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float sa = warp_sum(A[r]), sb = warp_sum(Bx[r]);
                if (ok[r]) {
                    sumA += sa;
                    sumB += sb;
                    if (lane == 0) {
                        atomicAdd(&S[(size_t)((b0 + r) * p.G + g) * 2], (double)(gm * sa));
                        atomicAdd(&S[(size_t)((b0 + r) * p.G + g) * 2 + 1], (double)(gm * sb));
                    }
                }
            }
        }
        if (lane == 0) {
            atomicAdd(&dgamma[c], sumB);
            atomicAdd(&dbeta[c], sumA);
        }
    }
}

template <typename OT, typename YT, typename DT, int PLANES>
__global__ void __launch_bounds__(kThreads, 3)
gn_bwd_pass2_kernel(const YT* __restrict__ y, const DT* __restrict__ dz, const float* __restrict__ mr,
                    const float* __restrict__ gamma, const double* __restrict__ S, OT* __restrict__ dy,
                    long long pstride, float* __restrict__ dbias, int C, int B, int T, int Tp, int G, float inv_n) {
    constexpr int R = RowsInFlight<YT, DT>::value;
    constexpr int planes = PLANES;
    extern __shared__ float sg_rows[];
    const int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (Tp + 8);
    constexpr bool multi = PLANES > 1;
    const int nseg_p = Tp >> 3;
    const bool shfl = nseg_p <= 32;
    if (multi && !shfl) srow_clear_halo(srow, Tp, lane);
    const int Cg = C / G;
    const int nchunk = (B + kChunkB - 1) / kChunkB;
    const int tasks = C * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const int c = task / nchunk, ch = task - c * nchunk;
        const int b_lo = ch * kChunkB, b_hi = min(B, b_lo + kChunkB);
        const int g = c / Cg;
        const float gm = __ldg(gamma + c);
        float db = 0.f;
        for (int b0 = b_lo; b0 < b_hi; b0 += R) {
            float c1[R], c2[R], c3[R];
            bool ok[R];
            long long row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < b_hi;
                const int bb = ok[r] ? b0 + r : b0;
                row[r] = ((long long)c * B + bb) * Tp;
                const float2 st = __ldg(mr2 + bb * G + g);
                const float m1 = (float)S[(size_t)(bb * G + g) * 2] * inv_n;
                const float m2 = (float)S[(size_t)(bb * G + g) * 2 + 1] * inv_n;
                c1[r] = st.y * gm;
                c2[r] = -st.y * st.y * m2;
                c3[r] = st.y * (st.x * st.y * m2 - m1);
            }
            auto finish = [&](int r, int seg, const F8& yv, const F8& zv, F8& o) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o.v[i] = fmaf(c1[r], zv.v[i], fmaf(c2[r], yv.v[i], c3[r]));
                if (seg * 8 + 8 > T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (seg * 8 + i >= T) o.v[i] = 0.f;
                }
                if (ok[r]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) db += o.v[i];
                }
            };
            if (shfl || !multi) {
                for (int seg = lane; seg < (shfl ? 32 : nseg_p); seg += 32) {
                    const bool live = seg * 8 < T;
                    typename RawOf<YT>::type yw[R];
                    typename RawOf<DT>::type zw[R];
                    if (live) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            yw[r] = load_raw(y + row[r] + seg * 8);
                            zw[r] = load_raw(dz + row[r] + seg * 8);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        F8 o;
                        if (live) {
                            finish(r, seg, cvt8(yw[r]), cvt8(zw[r]), o);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
                        }
                        if (multi) {
                            if (ok[r]) store_planes_shfl<OT, PLANES>(dy, row[r], pstride, o, T, nseg_p, lane);
                        } else if (ok[r] && seg < nseg_p) {
                            store8(dy + row[r] + seg * 8, o);
                        }
                    }
                }
            } else if (PLANES > 1) {
                // long rows with shifted planes: staged through shared memory, one row at a time
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    if (!ok[r]) continue;
                    for (int seg = lane; seg < nseg_p; seg += 32) {
                        F8 o;
                        if (seg * 8 < T) {
                            F8 yv = load8(y + row[r] + seg * 8), zv = load8(dz + row[r] + seg * 8);
                            finish(r, seg, yv, zv, o);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
                    }
                    __syncwarp();
                    store_row_planes(dy, row[r], planes, pstride, srow, T, Tp, lane);
                    __syncwarp();
                }
            }
        }
        db = warp_sum(db);
        if (lane == 0 && dbias != nullptr) atomicAdd(&dbias[c], db);
    }
}

// pass 2: dy = rstd * (gamma * dyh - mean_g(gamma dyh) - xhat * mean_g(gamma dyh xhat)) (or dyh without GroupNorm)
template <typename OT, typename RT, int ACT, bool POST>
__global__ void __launch_bounds__(kThreads)
gn_bwd_apply_kernel(BwdArgs p, const double* __restrict__ S, OT* __restrict__ dy, int planes, long long pstride,
                    float* __restrict__ dbias, float* __restrict__ dres, int dres_accumulate) {
    constexpr int R = kRowsInFlight;
    extern __shared__ float sg_rows[];
    const int lane = threadIdx.x & 31;
    float* srow = sg_rows + (threadIdx.x >> 5) * (p.Tp + 8);
    const bool multi = planes > 1;
    const int nseg_p = p.Tp >> 3;
    const bool shfl = multi && nseg_p <= 32;
    if (multi && !shfl) srow_clear_halo(srow, p.Tp, lane);
    const bool has_gn = p.mr != nullptr;
    const int Cg = has_gn ? p.C / p.G : 1;
    const int nchunk = (p.B + kChunkB - 1) / kChunkB;
    const int tasks = p.C * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(p.mr);
    const float inv_n = (float)p.inv_n;
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const int c = task / nchunk, ch = task - c * nchunk;
        const int b_lo = ch * kChunkB, b_hi = min(p.B, b_lo + kChunkB);
        const int g = has_gn ? c / Cg : 0;
        const float gm = has_gn ? __ldg(p.gamma + c) : 1.f, bt = has_gn ? __ldg(p.beta + c) : 0.f;
        float db = 0.f;
        if (nseg_p <= 32) {
            // rows of <= 256 elements: one segment per lane, the next row's loads in flight during the current row's math
            const bool live = lane * 8 < p.T;
            const long long row0 = (long long)c * p.B;
            SegIn cur, nxt;
            float2 st = make_float2(0.f, 1.f), st_n = st;
            double s1 = 0.0, s2 = 0.0, s1_n = 0.0, s2_n = 0.0;
            if (has_gn) {
                st = __ldg(mr2 + b_lo * p.G + g);
                s1 = S[(size_t)(b_lo * p.G + g) * 2];
                s2 = S[(size_t)(b_lo * p.G + g) * 2 + 1];
            }
            if (live) bwd_load<RT, POST>(p, row0 + b_lo, lane, cur);
#pragma unroll 1
            for (int b = b_lo; b < b_hi; ++b) {
                if (b + 1 < b_hi) {
                    if (live) bwd_load<RT, POST>(p, row0 + b + 1, lane, nxt);
                    if (has_gn) {
                        st_n = __ldg(mr2 + (b + 1) * p.G + g);
                        s1_n = S[(size_t)((b + 1) * p.G + g) * 2];
                        s2_n = S[(size_t)((b + 1) * p.G + g) * 2 + 1];
                    }
                }
                const long long row = row0 + b;
                F8 o, dpre;
                if (live) {
                    const float rstd = st.y, a = has_gn ? gm * st.y : 1.f, sh = has_gn ? bt - st.x * a : 0.f;
                    const float m1 = (float)s1 * inv_n, m2 = (float)s2 * inv_n;
                    F8 dyh, xhat;
                    bwd_compute<ACT, POST>(p, cur, lane, a, sh, st.x, rstd, dyh, xhat, dpre);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float v = dyh.v[i];
                        if (has_gn) v = rstd * (gm * v - m1 - xhat.v[i] * m2);
                        o.v[i] = v;
                    }
                    if (has_gn && lane * 8 + 8 > p.T) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (lane * 8 + i >= p.T) o.v[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) db += o.v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { o.v[i] = 0.f; dpre.v[i] = 0.f; }
                }
                if (dres != nullptr && lane < nseg_p) {
                    float* dr = dres + row * p.Tp + lane * 8;
                    if (dres_accumulate) {
                        F8 old = load8(dr);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dpre.v[i] += old.v[i];
                    }
                    store8(dr, dpre);
                }
                if (multi) store_planes_shfl_n(dy, row * p.Tp, planes, pstride, o, p.T, nseg_p, lane);
                else if (lane < nseg_p) store8(dy + row * p.Tp + lane * 8, o);
                cur = nxt;
                st = st_n;
                s1 = s1_n;
                s2 = s2_n;
            }
            db = warp_sum(db);
            if (lane == 0 && dbias != nullptr) atomicAdd(&dbias[c], db);
            continue;
        }
        for (int b0 = b_lo; b0 < b_hi; b0 += R) {
            float a[R], sh[R], mean[R], rstd[R], m1[R], m2[R];
            bool ok[R];
            long long row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < b_hi;
                const int bb = ok[r] ? b0 + r : b0;
                row[r] = (long long)c * p.B + bb;
                a[r] = 1.f; sh[r] = 0.f; mean[r] = 0.f; rstd[r] = 1.f; m1[r] = 0.f; m2[r] = 0.f;
                if (has_gn) {
                    const float2 st = __ldg(mr2 + bb * p.G + g);
                    mean[r] = st.x;
                    rstd[r] = st.y;
                    a[r] = gm * st.y;
                    sh[r] = bt - st.x * a[r];
                    m1[r] = (float)S[(size_t)(bb * p.G + g) * 2] * inv_n;
                    m2[r] = (float)S[(size_t)(bb * p.G + g) * 2 + 1] * inv_n;
                }
            }
            // one segment of one row: gradient wrt the conv output (o) and wrt the residual input (dpre)
            auto finish = [&](int r, int seg, const SegIn& in, F8& o, F8& dpre) {
                F8 dyh, xhat;
                bwd_compute<ACT, POST>(p, in, seg, a[r], sh[r], mean[r], rstd[r], dyh, xhat, dpre);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = dyh.v[i];
                    if (has_gn) v = rstd[r] * (gm * v - m1[r] - xhat.v[i] * m2[r]);
                    o.v[i] = v;
                }
                if (has_gn && seg * 8 + 8 > p.T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (seg * 8 + i >= p.T) o.v[i] = 0.f;
                }
                if (ok[r]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) db += o.v[i];
                }
            };
            auto store_dres = [&](int r, int seg, F8& dpre) {
                if (dres != nullptr && ok[r]) {
                    float* dr = dres + row[r] * p.Tp + seg * 8;
                    if (dres_accumulate) {
                        F8 old = load8(dr);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dpre.v[i] += old.v[i];
                    }
                    store8(dr, dpre);
                }
            };
            if (shfl || !multi) {
                // rows of <= 256 elements with shifted planes (one segment per lane), or single-plane rows of any length
                for (int seg = lane; seg < (shfl ? 32 : nseg_p); seg += 32) {
                    SegIn in[R];
                    const bool live = seg * 8 < p.T;
                    if (live) {
#pragma unroll
                        for (int r = 0; r < R; ++r) bwd_load<RT, POST>(p, row[r], seg, in[r]);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        F8 o, dpre;
                        if (live) {
                            finish(r, seg, in[r], o, dpre);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { o.v[i] = 0.f; dpre.v[i] = 0.f; }
                        }
                        if (seg < nseg_p) store_dres(r, seg, dpre);
                        if (shfl) {
                            if (ok[r]) store_planes_shfl_n(dy, row[r] * p.Tp, planes, pstride, o, p.T, nseg_p, lane);
                        } else if (ok[r]) {
                            store8(dy + row[r] * p.Tp + seg * 8, o);
                        }
                    }
                }
            } else {
                // long rows with shifted planes: staged through shared memory, one row at a time
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    if (!ok[r]) continue;
                    for (int seg = lane; seg < nseg_p; seg += 32) {
                        F8 o, dpre;
                        if (seg * 8 < p.T) {
                            SegIn in;
                            bwd_load<RT, POST>(p, row[r], seg, in);
                            finish(r, seg, in, o, dpre);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { o.v[i] = 0.f; dpre.v[i] = 0.f; }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) srow[4 + seg * 8 + i] = o.v[i];
                        store_dres(r, seg, dpre);
                    }
                    __syncwarp();
                    store_row_planes(dy, row[r] * p.Tp, planes, pstride, srow, p.T, p.Tp, lane);
                    __syncwarp();
                }
            }
        }
        db = warp_sum(db);
        if (lane == 0 && dbias != nullptr) atomicAdd(&dbias[c], db);
    }
}

// ---------------------------------------------------------------------------------------------
// recon head forward: x_hat = tanh(GN(y)) in the external layout + loss sums (+ the per-row partial
// sums of the GroupNorm backward, so that the backward needs ONE pass over y / x instead of two)
// rowsums[row] = (sum gL, sum gL*xn, sum gM, sum gM*xn) with gL = loss'(d)(1 - xh^2), gM = 2 d (1 - xh^2),
// xn = (y - mean) rstd: the backward scales them by the upstream loss gradients (they are linear in them).
// ---------------------------------------------------------------------------------------------
template <typename YT, bool VEC, bool MSE>
__global__ void __launch_bounds__(kThreads)
recon_fwd_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ x, float* __restrict__ x_hat,
                 double* __restrict__ loss_sums, float4* __restrict__ rowsums, int N, int B, int T, int Tp, int G,
                 int loss_kind) {
    __shared__ double shm[2][32];
    const int lane = threadIdx.x & 31;
    const int Cg = N / G;
    const long long rows = (long long)N * B;
    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    double d0 = 0.0, d1 = 0.0;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);
        float2 st = *reinterpret_cast<const float2*>(mr + 2 * (b * G + n / Cg));
        const float mean = st.x, rstd = st.y;
        float a = gamma[n] * rstd, sh = beta[n] - mean * a;
        const float nm = -mean * rstd;                        // xn = y * rstd + nm
        const YT* yrow = y + row * Tp;
        long long xo = ((long long)b * N + n) * T;
        float l0 = 0.f, l1 = 0.f, aL = 0.f, bL = 0.f, aM = 0.f, bM = 0.f;
        for (int seg = lane; seg * 8 < T; seg += 32) {
            F8 yv = load8(yrow + seg * 8);
            F8 xv, xh;
            if (x != nullptr) xv = load8_ext<VEC>(x + xo, seg * 8, T);
            const bool full = seg * 8 + 8 <= T;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float h = tanh_fast(yv.v[i] * a + sh);
                xh.v[i] = h;
                if (x != nullptr) {
                    float d = h - xv.v[i];
                    if (!full && seg * 8 + i >= T) d = 0.f;     // tail of a ragged row contributes nothing
                    l1 = fmaf(d, d, l1);
                    if (!MSE) l0 += loss_term(loss_kind, d);
                    if (rowsums != nullptr) {
                        float om = fmaf(-h, h, 1.f);
                        float xn = fmaf(yv.v[i], rstd, nm);
                        float gm = d * om;                    // x2 folded in after the loop
                        aM += gm;
                        bM = fmaf(gm, xn, bM);
                        if (!MSE) {
                            float gl = loss_grad(loss_kind, d) * om;
                            aL += gl;
                            bL = fmaf(gl, xn, bL);
                        }
                    }
                }
            }
            if (x_hat != nullptr) store8_ext<VEC>(x_hat + xo, seg * 8, T, xh);
        }
        if (rowsums != nullptr) {
            aM = 2.f * warp_sum(aM);
            bM = 2.f * warp_sum(bM);
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

// Fast paths of the training step: x present, T % 8 == 0 (every 8-element segment is entirely valid or entirely
// padding) and 16-byte aligned external rows.  No per-element predicates and 32-bit row arithmetic: the generic
// kernel above spends ~190 of its ~370 instructions per row on index math and predication and is issue-bound
// (ncu: issue slots 65 % busy at 58 % DRAM throughput).
// Fast-path mapping: one warp per CHANNEL, looping over the B samples of that channel two rows at a time.
//  * the rows (n, b = 0..B-1) are contiguous in y / dy, gamma/beta/group index are loaded once per channel and
//    the per-row index divisions disappear (ncu source counters before: 263 warp instructions per row, 143 of them
//    index math, predication and reductions; issue slots 72 % busy at 47 % of the DRAM bandwidth);
//  * kRowsInFlight rows are loaded before the first dependent instruction (one 200-element row is only 1.2 KB of
//    loads per warp: too little to cover the HBM latency at the occupancy these kernels reach);
//  * dbias[n] is a register sum over the channel's rows: no atomics, no memset.
// A lane-per-row mapping (each lane walking its own row) was measured 2-5x slower: 32 different cache lines per
// load instruction times ~48 resident warps overflow L1 before a line is consumed.

// The target x is either the external fp32 tensor [B][N][T] (XT = float) or - XT = 16-bit - the packed operand of the
// first encoder conv, [N][B][Tp] in the operand format: the same values the encoder consumes, in the layout of y, at half
// the bytes (fp16 mode: |x| <= 0.7 carries an absolute rounding error <= 2.4e-4).
template <typename XT>
__device__ __forceinline__ const XT* x_row(const XT* x, int n, int bb, int B, int T, int Tp, size_t xstride) {
    if (sizeof(XT) == 2) return x + ((size_t)n * B + bb) * Tp;
    return x + (size_t)n * T + bb * xstride;
}
__device__ __forceinline__ F8 load8_x(const float* p) {
    F8 r;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ F8 load8_x(const __nv_bfloat16* p) { return load8(p); }

template <typename YT, typename XT, bool MSE, bool ROWSUMS, bool XHAT>
__global__ void __launch_bounds__(kThreads, 4)
recon_fwd_fast_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const XT* __restrict__ x, float* __restrict__ x_hat,
                      double* __restrict__ loss_sums, float4* __restrict__ rowsums, int N, int B, int T, int Tp, int G,
                      int loss_kind) {
    constexpr int R = kRowsInFlight;
    __shared__ double shm[2][32];
    const int lane = threadIdx.x & 31;
    const int Cg = N / G;
    const int wstride = gridDim.x * kWarpsPerBlock;
    const int nseg = T >> 3;
    const size_t xstride = (size_t)N * T;                    // x / x_hat: distance between samples of one channel
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    double d0 = 0.0, d1 = 0.0;
    for (int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += wstride) {
        const float gam = __ldg(gamma + n), bet = __ldg(beta + n);
        const int g = n / Cg;
        const YT* ych = y + (size_t)n * B * Tp;
        float* hch = XHAT ? x_hat + (size_t)n * T : nullptr;
        float s0 = 0.f, s1 = 0.f;                            // fp32 partials of this channel, flushed to fp64 per channel
        for (int b0 = 0; b0 < B; b0 += R) {
            float a[R], sh[R], rstd[R], nm[R];
            bool ok[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < B;
                const float2 st = __ldg(mr2 + (ok[r] ? b0 + r : b0) * G + g);
                rstd[r] = st.y;
                a[r] = gam * st.y;
                sh[r] = bet - st.x * a[r];
                nm[r] = -st.x * st.y;
            }
            float l0[R], l1[R], aL[R], bL[R], aM[R], bM[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { l0[r] = l1[r] = aL[r] = bL[r] = aM[r] = bM[r] = 0.f; }
            for (int seg = lane; seg < nseg; seg += 32) {
                F8 yv[R], xw[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {                // every load of every row first
                    const int bb = ok[r] ? b0 + r : b0;
                    yv[r] = load8(ych + (size_t)bb * Tp + seg * 8);
                    xw[r] = load8_x(x_row(x, n, bb, B, T, Tp, xstride) + seg * 8);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float* xv = xw[r].v;
                    float h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        h[i] = tanh_fast(fmaf(yv[r].v[i], a[r], sh[r]));
                        const float d = h[i] - xv[i];
                        l1[r] = fmaf(d, d, l1[r]);
                        if (!MSE) l0[r] += loss_term(loss_kind, d);
                        if (ROWSUMS) {
                            const float om = fmaf(-h[i], h[i], 1.f);
                            const float xn = fmaf(yv[r].v[i], rstd[r], nm[r]);
                            const float gm = d * om;
                            aM[r] += gm;
                            bM[r] = fmaf(gm, xn, bM[r]);
                            if (!MSE) {
                                const float gl = loss_grad(loss_kind, d) * om;
                                aL[r] += gl;
                                bL[r] = fmaf(gl, xn, bL[r]);
                            }
                        }
                    }
                    if (XHAT && ok[r]) {
                        float* hp = hch + (b0 + r) * xstride + seg * 8;
                        *reinterpret_cast<float4*>(hp) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4*>(hp + 4) = make_float4(h[4], h[5], h[6], h[7]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (ROWSUMS) {
                    float am = 2.f * warp_sum(aM[r]), bm = 2.f * warp_sum(bM[r]);
                    float al = am, bl = bm;
                    if (!MSE) { al = warp_sum(aL[r]); bl = warp_sum(bL[r]); }
                    if (lane == 0 && ok[r]) rowsums[(size_t)n * B + b0 + r] = make_float4(al, bl, am, bm);
                }
                if (ok[r]) {
                    s0 += MSE ? l1[r] : l0[r];
                    s1 += l1[r];
                }
            }
        }
        d0 += (double)s0;
        d1 += (double)s1;
    }
    double t0 = block_sum(d0, shm[0]);
    double t1 = block_sum(d1, shm[1]);
    if (threadIdx.x == 0) {
        atomicAdd(&loss_sums[0], t0);
        atomicAdd(&loss_sums[1], t1);
    }
}

template <typename YT, typename XT, typename OT, bool MSE>
__global__ void __launch_bounds__(kThreads, 4)
recon_bwd_apply_fast_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const XT* __restrict__ x, const float* __restrict__ scal,
                            const double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias, int N, int B,
                            int T, int Tp, int G, int loss_kind, float inv_n) {
    constexpr int R = kRowsInFlight;
    const int lane = threadIdx.x & 31;
    const int Cg = N / G;
    const int wstride = gridDim.x * kWarpsPerBlock;
    const int nseg = T >> 3, nseg_p = Tp >> 3;
    const size_t xstride = (size_t)N * T;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    const float ga = scal[0], gm = scal[1];
    const float g2 = 2.f * (ga + gm);
    for (int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += wstride) {
        const float gam = __ldg(gamma + n), bet = __ldg(beta + n);
        const int g = n / Cg;
        const YT* ych = y + (size_t)n * B * Tp;
        OT* dch = dy + (size_t)n * B * Tp;
        float db = 0.f;
        for (int b0 = 0; b0 < B; b0 += R) {
            float a[R], sh[R], c1[R], c2[R], c3[R];
            bool ok[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = b0 + r < B;
                const int bb = ok[r] ? b0 + r : b0;
                const float2 st = __ldg(mr2 + bb * G + g);
                const float mean = st.x, rstd = st.y;
                a[r] = gam * rstd;
                sh[r] = bet - mean * a[r];
                const float m1 = (float)S[(size_t)(bb * G + g) * 2] * inv_n;
                const float m2 = (float)S[(size_t)(bb * G + g) * 2 + 1] * inv_n;
                c1[r] = rstd * gam;
                c2[r] = -rstd * rstd * m2;
                c3[r] = rstd * (mean * rstd * m2 - m1);
            }
            for (int seg = lane; seg < nseg_p; seg += 32) {
                if (seg < nseg) {
                    F8 yv[R], xw[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int bb = ok[r] ? b0 + r : b0;
                        yv[r] = load8(ych + (size_t)bb * Tp + seg * 8);
                        xw[r] = load8_x(x_row(x, n, bb, B, T, Tp, xstride) + seg * 8);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float* xv = xw[r].v;
                        F8 o;
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float h = tanh_fast(fmaf(yv[r].v[i], a[r], sh[r]));
                            const float d = h - xv[i];
                            const float gg = (MSE ? g2 * d : ga * loss_grad(loss_kind, d) + gm * 2.f * d) * fmaf(-h, h, 1.f);
                            o.v[i] = fmaf(c1[r], gg, fmaf(c2[r], yv[r].v[i], c3[r]));
                            acc += o.v[i];
                        }
                        if (ok[r]) {
                            store8(dch + (size_t)(b0 + r) * Tp + seg * 8, o);
                            db += acc;
                        }
                    }
                } else {
                    F8 o;
#pragma unroll
                    for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (ok[r]) store8(dch + (size_t)(b0 + r) * Tp + seg * 8, o);
                }
            }
        }
        db = warp_sum(db);
        if (lane == 0) dbias[n] = db;
    }
}

// ---------------------------------------------------------------------------------------------
// Short rows: Tp == 8 (T = 1 .. 8; the static configuration, Dim2 = 1, is T = 1 on 10^6 nodes x 512 samples).
// A warp per row leaves 31 of 32 lanes idle there (measured: 207 + 204 + 67 ms of a 556 ms step in recon_fwd, recon_bwd
// and pack_input at N = 10^6, B = 512).  Mapping of the short-row kernels: ONE THREAD PER ROW.  A block walks tiles of
// 32 channels x 32 samples; warp w takes channels n0 + w, n0 + w + 8, ... and lane l the sample b0 + l, so the 32 rows a
// warp touches per channel are 512 (16-bit) / 1024 (fp32) contiguous bytes of y / dy / the row sums.  The external fp32
// tensors ([B][N][T]: x, x_hat) are contiguous along (n, t) instead: their tile goes through shared memory, read and
// written in 32*T-float runs per sample.  No warp reductions per row (a row is one thread); dbias is one warp_sum per
// (channel, 32 samples).
// ---------------------------------------------------------------------------------------------
constexpr int kShortTile = 32;
constexpr int kShortPitch = kShortTile * 8 + 1;        // floats per sample in the staged tile (+1: conflict-free columns)
constexpr int kShortRows = kShortTile / kWarpsPerBlock;  // channels per warp and tile

// x[b0 .. b0+32)[n0 .. n0+32)[0 .. T) -> xs[sample][channel * T + t]
__device__ __forceinline__ void short_stage_in(const float* __restrict__ x, float (*xs)[kShortPitch], int n0, int b0, int N,
                                               int B, int T) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run = min(kShortTile, N - n0) * T;
    for (int bl = warp; bl < kShortTile && b0 + bl < B; bl += kWarpsPerBlock) {
        const float* src = x + ((size_t)(b0 + bl) * N + n0) * T;
        for (int i = lane; i < run; i += 32) xs[bl][i] = __ldg(src + i);
    }
}
__device__ __forceinline__ void short_stage_out(float* __restrict__ x, float (*xs)[kShortPitch], int n0, int b0, int N,
                                                int B, int T) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run = min(kShortTile, N - n0) * T;
    for (int bl = warp; bl < kShortTile && b0 + bl < B; bl += kWarpsPerBlock) {
        float* dst = x + ((size_t)(b0 + bl) * N + n0) * T;
        for (int i = lane; i < run; i += 32) dst[i] = xs[bl][i];
    }
}

template <typename OT>
__global__ void __launch_bounds__(kThreads)
pack_input_short_kernel(const float* __restrict__ x, OT* __restrict__ out, int B, int N, int T) {
    __shared__ float xs[kShortTile][kShortPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_b = (B + kShortTile - 1) / kShortTile;
    const long long tiles = (long long)((N + kShortTile - 1) / kShortTile) * tiles_b;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n0 = (int)(tile / tiles_b) * kShortTile, b0 = (int)(tile % tiles_b) * kShortTile;
        __syncthreads();
        short_stage_in(x, xs, n0, b0, N, B, T);
        __syncthreads();
        const int b = b0 + lane;
#pragma unroll
        for (int j = 0; j < kShortRows; ++j) {
            const int nl = warp + j * kWarpsPerBlock, n = n0 + nl;
            if (n < N && b < B) {
                F8 r;
#pragma unroll
                for (int i = 0; i < 8; ++i) r.v[i] = i < T ? xs[lane][nl * T + i] : 0.f;
                store8(out + ((size_t)n * B + b) * 8, r);
            }
        }
    }
}

// forward of the reconstruction head on short rows; x may be NULL (no loss: inference), XT = 16-bit: the packed operand
template <typename YT, typename XT, bool MSE, bool ROWSUMS, bool XHAT>
__global__ void __launch_bounds__(kThreads)
recon_fwd_short_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const XT* __restrict__ x, float* __restrict__ x_hat,
                       double* __restrict__ loss_sums, float4* __restrict__ rowsums, int N, int B, int T, int G,
                       int loss_kind) {
    constexpr bool kStageX = sizeof(XT) == 4;
    __shared__ float xs[(kStageX || XHAT) ? kShortTile : 1][kShortPitch];
    __shared__ double shm[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cg = N / G;
    const int tiles_b = (B + kShortTile - 1) / kShortTile;
    const long long tiles = (long long)((N + kShortTile - 1) / kShortTile) * tiles_b;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    const bool has_x = x != nullptr;
    double d0 = 0.0, d1 = 0.0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n0 = (int)(tile / tiles_b) * kShortTile, b0 = (int)(tile % tiles_b) * kShortTile;
        const int b = b0 + lane;
        typename RawOf<YT>::type yw[kShortRows];              // unconverted: nothing depends on the loads before the barrier
        typename RawOf<XT>::type xr[kShortRows];
        bool ok[kShortRows];
#pragma unroll
        for (int j = 0; j < kShortRows; ++j) {                // every global load of the tile first: the y rows are in
            const int n = n0 + warp + j * kWarpsPerBlock;     // flight while the x tile is staged
            ok[j] = n < N && b < B;
            const size_t row = ok[j] ? (size_t)n * B + b : 0;
            yw[j] = load_raw(y + row * 8);
            if (!kStageX && has_x) xr[j] = load_raw(x + row * 8);
        }
        if (kStageX || XHAT) __syncthreads();                 // the previous tile's readers / writers are done
        if (kStageX && has_x) {
            short_stage_in(reinterpret_cast<const float*>(x), xs, n0, b0, N, B, T);
            __syncthreads();
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < kShortRows; ++j) {
            const int nl = warp + j * kWarpsPerBlock, n = ok[j] ? n0 + nl : 0;
            const float gam = __ldg(gamma + n), bet = __ldg(beta + n);
            const float2 st = __ldg(mr2 + (ok[j] ? b : 0) * G + n / Cg);
            const float a = gam * st.y, sh = bet - st.x * a, nm = -st.x * st.y;
            float l0 = 0.f, l1 = 0.f, aL = 0.f, bL = 0.f, aM = 0.f, bM = 0.f;
            const F8 yv = cvt8(yw[j]);
            F8 xw;
            if (!kStageX && has_x) xw = cvt8(xr[j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < T) {
                    const float h = tanh_fast(fmaf(yv.v[i], a, sh));
                    if (has_x) {
                        const float xv = kStageX ? xs[lane][nl * T + i] : xw.v[i];
                        const float d = h - xv;
                        l1 = fmaf(d, d, l1);
                        if (!MSE) l0 += loss_term(loss_kind, d);
                        if (ROWSUMS) {
                            const float om = fmaf(-h, h, 1.f);
                            const float xn = fmaf(yv.v[i], st.y, nm);
                            const float gm = d * om;
                            aM += gm;
                            bM = fmaf(gm, xn, bM);
                            if (!MSE) {
                                const float gl = loss_grad(loss_kind, d) * om;
                                aL += gl;
                                bL = fmaf(gl, xn, bL);
                            }
                        }
                    }
                    if (XHAT) xs[lane][nl * T + i] = h;       // this thread's own slot of the tile
                }
            }
            if (ok[j]) {
                if (ROWSUMS) {
                    const float am = 2.f * aM, bm = 2.f * bM;
                    rowsums[(size_t)n * B + b] = make_float4(MSE ? am : aL, MSE ? bm : bL, am, bm);
                }
                s0 += MSE ? l1 : l0;
                s1 += l1;
            }
        }
        d0 += (double)s0;
        d1 += (double)s1;
        if (XHAT) {
            __syncthreads();
            short_stage_out(x_hat, xs, n0, b0, N, B, T);
        }
    }
    if (has_x) {
        double t0 = block_sum(d0, shm[0]);
        double t1 = block_sum(d1, shm[1]);
        if (threadIdx.x == 0) {
            atomicAdd(&loss_sums[0], t0);
            atomicAdd(&loss_sums[1], t1);
        }
    }
}

// one-pass backward (step 2) on short rows: dy = c1 * g + c2 * y + c3 per element, zero padding, dbias accumulated
template <typename YT, typename XT, typename OT, bool MSE>
__global__ void __launch_bounds__(kThreads)
recon_bwd_apply_short_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const XT* __restrict__ x, const float* __restrict__ scal,
                             const double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias, int N, int B,
                             int T, int G, int loss_kind, float inv_n) {
    constexpr bool kStageX = sizeof(XT) == 4;
    __shared__ float xs[kStageX ? kShortTile : 1][kShortPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cg = N / G;
    const int tiles_b = (B + kShortTile - 1) / kShortTile;
    const long long tiles = (long long)((N + kShortTile - 1) / kShortTile) * tiles_b;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    const float ga = scal[0], gm = scal[1];
    const float g2 = 2.f * (ga + gm);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n0 = (int)(tile / tiles_b) * kShortTile, b0 = (int)(tile % tiles_b) * kShortTile;
        const int b = b0 + lane;
        typename RawOf<YT>::type yw[kShortRows];
        typename RawOf<XT>::type xr[kShortRows];
        bool ok[kShortRows];
#pragma unroll
        for (int j = 0; j < kShortRows; ++j) {                // in flight while the x tile is staged
            const int n = n0 + warp + j * kWarpsPerBlock;
            ok[j] = n < N && b < B;
            const size_t row = ok[j] ? (size_t)n * B + b : 0;
            yw[j] = load_raw(y + row * 8);
            if (!kStageX) xr[j] = load_raw(x + row * 8);
        }
        if (kStageX) {
            __syncthreads();
            short_stage_in(reinterpret_cast<const float*>(x), xs, n0, b0, N, B, T);
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < kShortRows; ++j) {
            const int nl = warp + j * kWarpsPerBlock, n = n0 + nl;          // warp-uniform
            if (n >= N) break;
            const int bb = ok[j] ? b : 0, g = n / Cg;
            const float gam = __ldg(gamma + n), bet = __ldg(beta + n);
            const float2 st = __ldg(mr2 + bb * G + g);
            const float mean = st.x, rstd = st.y;
            const float a = gam * rstd, sh = bet - mean * a;
            const float m1 = (float)S[(size_t)(bb * G + g) * 2] * inv_n;
            const float m2 = (float)S[(size_t)(bb * G + g) * 2 + 1] * inv_n;
            const float c1 = rstd * gam, c2 = -rstd * rstd * m2, c3 = rstd * (mean * rstd * m2 - m1);
            F8 o;
            float acc = 0.f;
            const F8 yv = cvt8(yw[j]);
            F8 xw;
            if (!kStageX) xw = cvt8(xr[j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o.v[i] = 0.f;
                if (i < T) {
                    const float h = tanh_fast(fmaf(yv.v[i], a, sh));
                    const float d = h - (kStageX ? xs[lane][nl * T + i] : xw.v[i]);
                    const float gg = (MSE ? g2 * d : ga * loss_grad(loss_kind, d) + gm * 2.f * d) * fmaf(-h, h, 1.f);
                    o.v[i] = fmaf(c1, gg, fmaf(c2, yv.v[i], c3));
                    acc += o.v[i];
                }
            }
            if (ok[j]) store8(dy + ((size_t)n * B + b) * 8, o);
            acc = warp_sum(ok[j] ? acc : 0.f);
            if (lane == 0) atomicAdd(dbias + n, acc);
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
template <typename YT, typename OT, bool VEC, bool MSE>
__global__ void __launch_bounds__(kThreads)
recon_bwd_apply_kernel(const YT* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ x, const float* __restrict__ scal,
                       const double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias, int N, int B, int T,
                       int Tp, int G, int loss_kind, double inv_n) {
    const int lane = threadIdx.x & 31;
    const int Cg = N / G;
    const long long rows = (long long)N * B;
    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    const float ga = scal[0], gm = scal[1];
    const float g2 = 2.f * (ga + gm);      // MSE: ga * 2d + gm * 2d
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int n = (int)(row / B), b = (int)(row % B);
        int g = n / Cg;
        float2 st = *reinterpret_cast<const float2*>(mr + 2 * (b * G + g));
        const float mean = st.x, rstd = st.y;
        float gam = gamma[n];
        float a = gam * rstd, sh = beta[n] - mean * a;
        // dy = rstd*gam*gg - rstd*m1 - xn*rstd*m2 with xn = y*rstd - mean*rstd  ->  c1*gg + c2*y + c3
        const float m1 = (float)(S[(size_t)(b * G + g) * 2] * inv_n);
        const float m2 = (float)(S[(size_t)(b * G + g) * 2 + 1] * inv_n);
        const float c1 = rstd * gam, c2 = -rstd * rstd * m2, c3 = rstd * (mean * rstd * m2 - m1);
        const YT* yrow = y + row * Tp;
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
                    float gg = (MSE ? g2 * d : ga * loss_grad(loss_kind, d) + gm * 2.f * d) * fmaf(-h, h, 1.f);
                    o.v[i] = fmaf(c1, gg, fmaf(c2, yv.v[i], c3));
                }
                if (seg * 8 + 8 > T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (seg * 8 + i >= T) o.v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) db += o.v[i];
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

// two-pass recon backward (used when a gradient wrt x_hat itself arrives, or without the forward's row sums)
struct ReconBwdArgs {
    const void* y;
    const float* mr;
    const float* gamma;
    const float* beta;
    const float* x;
    const float* ext;
    int loss_kind, N, B, T, Tp, G;
    double inv_n;
};

template <typename YT, typename OT, bool REDUCE>
__global__ void __launch_bounds__(kThreads)
recon_bwd_kernel(ReconBwdArgs p, const float* __restrict__ scal, float* __restrict__ dgamma, float* __restrict__ dbeta,
                 double* __restrict__ S, OT* __restrict__ dy, float* __restrict__ dbias) {
    const float ga = scal[0], gmse = scal[1];
    const int lane = threadIdx.x & 31;
    const int Cg = p.N / p.G;
    const long long rows = (long long)p.N * p.B, wstride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += wstride) {
        int c = (int)(row / p.B), b = (int)(row % p.B);
        int g = c / Cg;
        float2 st = *reinterpret_cast<const float2*>(p.mr + 2 * (b * p.G + g));
        const float mean = st.x, rstd = st.y;
        float gm = p.gamma[c];
        float a = gm * rstd, sh = p.beta[c] - mean * a;
        float m1 = 0.f, m2 = 0.f;
        if (!REDUCE) {
            m1 = (float)(S[(size_t)(b * p.G + g) * 2] * p.inv_n);
            m2 = (float)(S[(size_t)(b * p.G + g) * 2 + 1] * p.inv_n);
        }
        float A = 0.f, Bx = 0.f, db = 0.f;
        for (int seg = lane; seg < p.Tp / 8; seg += 32) {
            F8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
            if (seg * 8 < p.T) {
                F8 yv = load8(reinterpret_cast<const YT*>(p.y) + row * p.Tp + seg * 8);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int t = seg * 8 + i;
                    if (t < p.T) {
                        float h = tanh_fast(yv.v[i] * a + sh);
                        long long xi = ((long long)b * p.N + c) * p.T + t;
                        float d = 0.f;
                        if (p.x != nullptr) {
                            float diff = h - __ldg(p.x + xi);
                            d = ga * loss_grad(p.loss_kind, diff) + gmse * 2.f * diff;
                        }
                        if (p.ext != nullptr) d += __ldg(p.ext + xi);
                        float gg = d * (1.f - h * h);
                        float xn = (yv.v[i] - mean) * rstd;
                        if (REDUCE) {
                            A += gg;
                            Bx += gg * xn;
                        } else {
                            float v = rstd * (gm * gg - m1 - xn * m2);
                            o.v[i] = v;
                            db += v;
                        }
                    }
                }
            }
            if (!REDUCE) store8(dy + row * p.Tp + seg * 8, o);
        }
        if (REDUCE) {
            A = warp_sum(A);
            Bx = warp_sum(Bx);
            if (lane == 0) {
                atomicAdd(&dgamma[c], Bx);
                atomicAdd(&dbeta[c], A);
                atomicAdd(&S[(size_t)(b * p.G + g) * 2], (double)(gm * A));
                atomicAdd(&S[(size_t)(b * p.G + g) * 2 + 1], (double)(gm * Bx));
            }
        } else {
            db = warp_sum(db);
            if (lane == 0) atomicAdd(&dbias[c], db);
        }
    }
}

// g_loss / g_mse are device scalars produced by autograd (may be NULL); fold them with inv_numel on device.
__global__ void recon_scalars_kernel(const float* g_loss, const float* g_mse, float inv_numel, float* out2) {
    out2[0] = g_loss ? g_loss[0] * inv_numel : 0.f;
    out2[1] = g_mse ? g_mse[0] * inv_numel : 0.f;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm / activation kernels on short rows (Tp == 8, static fields): one thread per row, like the recon head above.
// A warp takes a (slice of <= 8 channels of one group, 32 consecutive samples) task: lane = sample, so every access of a
// channel is 32 contiguous rows, the per-channel sums are one warp_sum + one atomic per (channel, 32 samples) and the
// per-(sample, group) sums of the backward stay in the lane's registers over the slice (2 double atomics per lane and
// task instead of 2 per row).  The layers these run on are small (C <= 5120 channels), so dtype, activation and plane
// count are RUN-TIME arguments (uniform branches): three kernels instead of another ~150 template instantiations.
// ---------------------------------------------------------------------------------------------
constexpr int kShortSlice = 8;

__device__ __forceinline__ F8 load8_rt(const void* p, bool is16, size_t elem) {
    return is16 ? load8(reinterpret_cast<const __nv_bfloat16*>(p) + elem) : load8(reinterpret_cast<const float*>(p) + elem);
}
__device__ __forceinline__ void store8_rt(void* p, bool is16, size_t elem, const F8& r) {
    if (is16) store8(reinterpret_cast<__nv_bfloat16*>(p) + elem, r);
    else store8(reinterpret_cast<float*>(p) + elem, r);
}
// o holds zeros for t >= T: plane pl is the row shifted by pl - PLANES / 2, zero outside [0, T)
template <int PLANES>
__device__ __forceinline__ void store_planes_short(void* base, bool is16, size_t row_off, long long pstride, const F8& o, int T) {
#pragma unroll
    for (int pl = 0; pl < PLANES; ++pl) {
        F8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = i + pl - PLANES / 2;
            r.v[i] = (j >= 0 && j < 8 && i < T) ? o.v[j] : 0.f;
        }
        store8_rt(base, is16, (size_t)pl * pstride + row_off, r);
    }
}
__device__ __forceinline__ void store_planes_short_n(void* base, bool is16, size_t row_off, int planes, long long pstride,
                                                     const F8& o, int T) {
    if (planes == 5) store_planes_short<5>(base, is16, row_off, pstride, o, T);
    else if (planes == 3) store_planes_short<3>(base, is16, row_off, pstride, o, T);
    else store8_rt(base, is16, row_off, o);
}

struct ShortTask {
    int c_lo, c_hi, g, b;
    bool ok;
};
// task -> (channel slice inside one group, sample of this lane); Cg = channels per group (C without GroupNorm)
__device__ __forceinline__ ShortTask short_task(int task, int nchunk, int slices, int Cg, int B, int lane) {
    ShortTask t;
    const int cs = task / nchunk, ch = task - cs * nchunk;
    t.g = cs / slices;
    t.c_lo = t.g * Cg + (cs - t.g * slices) * kShortSlice;
    t.c_hi = min(t.c_lo + kShortSlice, (t.g + 1) * Cg);
    t.b = ch * 32 + lane;
    t.ok = t.b < B;
    if (!t.ok) t.b = B - 1;
    return t;
}

__global__ void __launch_bounds__(kThreads)
gn_act_fwd_short_kernel(const void* __restrict__ y, int y16, const float* __restrict__ mr, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const void* __restrict__ res, int r16, float res_scale, int act,
                        int post, void* __restrict__ out_op, int o16, int planes, long long pstride,
                        float* __restrict__ out_f32, int C, int B, int T, int G) {
    const int lane = threadIdx.x & 31;
    const bool has_gn = mr != nullptr;
    const int Cg = has_gn ? C / G : C, groups = has_gn ? G : 1;
    const int slices = (Cg + kShortSlice - 1) / kShortSlice, nchunk = (B + 31) / 32;
    const int tasks = groups * slices * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const ShortTask t = short_task(task, nchunk, slices, Cg, B, lane);
        float mean = 0.f, rstd = 1.f;
        if (has_gn) {
            const float2 st = __ldg(mr2 + t.b * G + t.g);
            mean = st.x;
            rstd = st.y;
        }
#pragma unroll 2
        for (int c = t.c_lo; c < t.c_hi; ++c) {
            const size_t row = ((size_t)c * B + t.b) * 8;
            const F8 yv = load8_rt(y, y16, row);
            F8 rv;
            if (res != nullptr) rv = load8_rt(res, r16, row);
            float a = 1.f, sh = 0.f;
            if (has_gn) {
                a = __ldg(gamma + c) * rstd;
                sh = __ldg(beta + c) - mean * a;
            }
            F8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o.v[i] = 0.f;
                if (i < T) {
                    float pre = res_scale * act_f(act, fmaf(yv.v[i], a, sh));
                    if (res != nullptr) pre += rv.v[i];
                    if (post) pre = gelu_f(pre);
                    o.v[i] = pre;
                }
            }
            if (t.ok) {
                if (out_f32 != nullptr) store8(out_f32 + row, o);
                if (out_op != nullptr) store_planes_short_n(out_op, o16, row, planes, pstride, o, T);
            }
        }
    }
}

// pass 1 of the GroupNorm backward (see gn_bwd_pass1_kernel): dout -> dz in place, dres, dgamma / dbeta / S
__global__ void __launch_bounds__(kThreads)
gn_bwd_pass1_short_kernel(Pass1Args p, int y16, int d16, int r16, int act, int post, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, double* __restrict__ S, float* __restrict__ dres, int dres_accumulate) {
    const int lane = threadIdx.x & 31;
    const int Cg = p.C / p.G;
    const int slices = (Cg + kShortSlice - 1) / kShortSlice, nchunk = (p.B + 31) / 32;
    const int tasks = p.G * slices * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(p.mr);
    const bool has_res = p.res != nullptr;
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const ShortTask t = short_task(task, nchunk, slices, Cg, p.B, lane);
        const float2 st = __ldg(mr2 + t.b * p.G + t.g);
        const float mean = st.x, rstd = st.y, nm = -st.x * st.y;
        float SA = 0.f, SB = 0.f;                             // gamma-weighted sums of this lane's sample over the slice
#pragma unroll 2
        for (int c = t.c_lo; c < t.c_hi; ++c) {
            const size_t row = ((size_t)c * p.B + t.b) * 8;
            const F8 yv = load8_rt(p.y, y16, row), dv = load8_rt(p.dout, d16, row);
            F8 rv;
            if (post && has_res) rv = load8_rt(p.res, r16, row);
            const float gm = __ldg(p.gamma + c);
            const float a = gm * rstd, sh = __ldg(p.beta + c) - mean * a;
            F8 dz, dp;
            float A = 0.f, Bx = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dz.v[i] = 0.f;
                dp.v[i] = 0.f;
                if (i < p.T) {
                    float val, dval;
                    act_both(act, fmaf(yv.v[i], a, sh), val, dval);
                    float d = dv.v[i];
                    if (post) d *= gelu_grad_f(fmaf(p.res_scale, val, has_res ? rv.v[i] : 0.f));
                    dp.v[i] = d;
                    dz.v[i] = p.res_scale * d * dval;
                    A += dz.v[i];
                    Bx = fmaf(dz.v[i], fmaf(yv.v[i], rstd, nm), Bx);
                }
            }
            if (t.ok) {
                store8_rt(p.dout, d16, row, dz);
                if (dres != nullptr) {
                    if (dres_accumulate) {
                        const F8 old = load8(dres + row);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dp.v[i] += old.v[i];
                    }
                    store8(dres + row, dp);
                }
            } else {
                A = 0.f;
                Bx = 0.f;
            }
            SA = fmaf(gm, A, SA);
            SB = fmaf(gm, Bx, SB);
            const float sa = warp_sum(A), sb = warp_sum(Bx);
            if (lane == 0) {
                atomicAdd(&dgamma[c], sb);
                atomicAdd(&dbeta[c], sa);
            }
        }
        if (t.ok) {
            atomicAdd(&S[(size_t)(t.b * p.G + t.g) * 2], (double)SA);
            atomicAdd(&S[(size_t)(t.b * p.G + t.g) * 2 + 1], (double)SB);
        }
    }
}

// pass 2: dy = c1 * dz + c2 * y + c3 (zero padding), shifted operand planes, dbias
__global__ void __launch_bounds__(kThreads)
gn_bwd_pass2_short_kernel(const void* __restrict__ y, int y16, const void* __restrict__ dz, int d16,
                          const float* __restrict__ mr, const float* __restrict__ gamma, const double* __restrict__ S,
                          void* __restrict__ dy, int o16, int planes, long long pstride, float* __restrict__ dbias, int C,
                          int B, int T, int G, float inv_n) {
    const int lane = threadIdx.x & 31;
    const int Cg = C / G;
    const int slices = (Cg + kShortSlice - 1) / kShortSlice, nchunk = (B + 31) / 32;
    const int tasks = G * slices * nchunk, wstride = gridDim.x * kWarpsPerBlock;
    const float2* mr2 = reinterpret_cast<const float2*>(mr);
    for (int task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < tasks; task += wstride) {
        const ShortTask t = short_task(task, nchunk, slices, Cg, B, lane);
        const float2 st = __ldg(mr2 + t.b * G + t.g);
        const float m1 = (float)S[(size_t)(t.b * G + t.g) * 2] * inv_n;
        const float m2 = (float)S[(size_t)(t.b * G + t.g) * 2 + 1] * inv_n;
        const float c2 = -st.y * st.y * m2, c3 = st.y * (st.x * st.y * m2 - m1);
#pragma unroll 2
        for (int c = t.c_lo; c < t.c_hi; ++c) {
            const size_t row = ((size_t)c * B + t.b) * 8;
            const F8 yv = load8_rt(y, y16, row), zv = load8_rt(dz, d16, row);
            const float c1 = st.y * __ldg(gamma + c);
            F8 o;
            float db = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o.v[i] = 0.f;
                if (i < T) {
                    o.v[i] = fmaf(c1, zv.v[i], fmaf(c2, yv.v[i], c3));
                    db += o.v[i];
                }
            }
            if (t.ok) store_planes_short_n(dy, o16, row, planes, pstride, o, T);
            db = warp_sum(t.ok ? db : 0.f);
            if (lane == 0 && dbias != nullptr) atomicAdd(&dbias[c], db);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch over the template parameters
// ---------------------------------------------------------------------------------------------
template <typename OT, typename RT, typename YT, int ACT, bool POST>
static void launch_fwd_t(const void* y, const float* mr, const float* gamma, const float* beta, const void* res,
                         float res_scale, void* out_op, int planes, long long pstride, float* out_f32, int C, int B, int T,
                         int Tp, int G, cudaStream_t st) {
    size_t sm = (planes > 1 && out_op && (Tp >> 3) > 32) ? sizeof(float) * kWarpsPerBlock * (Tp + 8) : 0;
    const long long tasks = (long long)C * cdiv(B, kChunkB);
#define SG_FP(PL) gn_act_fwd_kernel<OT, RT, YT, ACT, POST, PL><<<task_grid(gn_act_fwd_kernel<OT, RT, YT, ACT, POST, PL>, tasks, sm), kThreads, sm, st>>>( \
        (const YT*)y, mr, gamma, beta, (const RT*)res, res_scale, (OT*)out_op, pstride, out_f32, C, B, T, Tp, G)
    if (out_op == nullptr || planes == 1) SG_FP(1); else if (planes == 3) SG_FP(3); else SG_FP(5);
#undef SG_FP
}

template <typename OT, typename RT, typename YT>
static int launch_fwd(int act, int post, const void* y, const float* mr, const float* gamma, const float* beta,
                      const void* res, float res_scale, void* out_op, int planes, long long pstride, float* out_f32, int C,
                      int B, int T, int Tp, int G, cudaStream_t st) {
#define SG_F(ACT, POST) launch_fwd_t<OT, RT, YT, ACT, POST>(y, mr, gamma, beta, res, res_scale, out_op, planes, pstride, out_f32, C, B, T, Tp, G, st)
    if (act == SG_ACT_GELU) { if (post) SG_F(SG_ACT_GELU, true); else SG_F(SG_ACT_GELU, false); }
    else if (act == SG_ACT_TANH) { if (post) SG_F(SG_ACT_TANH, true); else SG_F(SG_ACT_TANH, false); }
    else { if (post) SG_F(SG_ACT_NONE, true); else SG_F(SG_ACT_NONE, false); }
#undef SG_F
    return check_launch("gn_act_fwd");
}

// layers without GroupNorm (ConvTranspose1d + GELU, plain convs): one pass, fp32 y and dout
template <typename OT, typename RT, int ACT, bool POST>
static void launch_bwd_plain_t(const BwdArgs& p, OT* dy, int planes, long long pstride, float* dbias, float* dres,
                               int dres_accumulate, double* ws, cudaStream_t st) {
    size_t sm = planes > 1 ? sizeof(float) * kWarpsPerBlock * (p.Tp + 8) : 0;
    int grid = persistent_grid((long long)p.C * cdiv(p.B, kChunkB));
    gn_bwd_apply_kernel<OT, RT, ACT, POST><<<grid, kThreads, sm, st>>>(p, ws, dy, planes, pstride, dbias, dres, dres_accumulate);
}

template <typename OT, typename RT>
static int launch_bwd_plain(int act, int post, const BwdArgs& p, OT* dy, int planes, long long pstride, float* dbias,
                            float* dres, int dres_accumulate, double* ws, cudaStream_t st) {
#define SG_B(ACT, POST) launch_bwd_plain_t<OT, RT, ACT, POST>(p, dy, planes, pstride, dbias, dres, dres_accumulate, ws, st)
    if (act == SG_ACT_GELU) { if (post) SG_B(SG_ACT_GELU, true); else SG_B(SG_ACT_GELU, false); }
    else if (act == SG_ACT_TANH) { if (post) SG_B(SG_ACT_TANH, true); else SG_B(SG_ACT_TANH, false); }
    else { if (post) SG_B(SG_ACT_NONE, true); else SG_B(SG_ACT_NONE, false); }
#undef SG_B
    return check_launch("gn_act_bwd");
}

// GroupNorm layers: pass 1 (dout -> dz in place, reductions, dres) + pass 2 (dy planes, dbias)
template <typename OT, typename RT, typename YT, typename DT>
static int launch_bwd_gn(int act, int post, const Pass1Args& p, OT* dy, int planes, long long pstride, float* dgamma,
                         float* dbeta, float* dbias, float* dres, int dres_accumulate, double* ws, cudaStream_t st) {
    const long long tasks = (long long)p.C * cdiv(p.B, kChunkB);
#define SG_P1(ACT, POST) gn_bwd_pass1_kernel<RT, YT, DT, ACT, POST><<<task_grid(gn_bwd_pass1_kernel<RT, YT, DT, ACT, POST>, tasks, 0), kThreads, 0, st>>>(p, dgamma, dbeta, ws, dres, dres_accumulate)
    if (act == SG_ACT_GELU) { if (post) SG_P1(SG_ACT_GELU, true); else SG_P1(SG_ACT_GELU, false); }
    else if (act == SG_ACT_TANH) { if (post) SG_P1(SG_ACT_TANH, true); else SG_P1(SG_ACT_TANH, false); }
    else { if (post) SG_P1(SG_ACT_NONE, true); else SG_P1(SG_ACT_NONE, false); }
#undef SG_P1
    const size_t sm = (planes > 1 && (p.Tp >> 3) > 32) ? sizeof(float) * kWarpsPerBlock * (p.Tp + 8) : 0;
    const float inv_n = (float)(1.0 / ((double)(p.C / p.G) * p.T));
#define SG_P2(PL) gn_bwd_pass2_kernel<OT, YT, DT, PL><<<task_grid(gn_bwd_pass2_kernel<OT, YT, DT, PL>, tasks, sm), kThreads, sm, st>>>((const YT*)p.y, (const DT*)p.dout, p.mr, p.gamma, ws, \
                                                                                dy, pstride, dbias, p.C, p.B, p.T, p.Tp, p.G, inv_n)
    if (planes == 1) SG_P2(1); else if (planes == 3) SG_P2(3); else SG_P2(5);
#undef SG_P2
    return check_launch("gn_act_bwd");
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_pack_input(const void* xv, int x_dtype, void* out, int B, int N, int T, int Tp, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CHECK_OP16(x_dtype);
    SG_REQUIRE(Tp % 8 == 0 && Tp >= T, "pack_input: bad Tp=%d for T=%d", Tp, T);
    long long rows = (long long)N * B;
    int grid = persistent_grid(rows) * 2;
    cudaStream_t st = as_stream(stream);
    if (is_op16(x_dtype)) {
        SG_REQUIRE(x_dtype == dtype && T % 8 == 0 && Tp == T && aligned16(xv) && aligned16(out),
                   "pack_input: a 16-bit source needs the same operand format, T %% 8 == 0 and 16-byte alignment");
        pack_input16_kernel<<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)xv, (__nv_bfloat16*)out, B, N, T);
        return check_launch("pack_input");
    }
    const float* x = (const float*)xv;
    if (Tp == 8 && aligned16(out)) {                         // short rows: one thread per row (static fields)
        if (is_op16(dtype))
            pack_input_short_kernel<__nv_bfloat16><<<resident_grid(pack_input_short_kernel<__nv_bfloat16>, short_tiles(N, B)), kThreads, 0, st>>>(
                x, (__nv_bfloat16*)out, B, N, T);
        else
            pack_input_short_kernel<float><<<resident_grid(pack_input_short_kernel<float>, short_tiles(N, B)), kThreads, 0, st>>>(x, (float*)out, B, N, T);
        return check_launch("pack_input");
    }
    bool vec = (T % 4 == 0) && aligned16(x);
#define SG_PACK(OT, VEC) pack_input_kernel<OT, VEC><<<grid, kThreads, 0, st>>>(x, (OT*)out, B, N, T, Tp)
    if (is_op16(dtype)) { if (vec) SG_PACK(__nv_bfloat16, true); else SG_PACK(__nv_bfloat16, false); }
    else                  { if (vec) SG_PACK(float, true); else SG_PACK(float, false); }
#undef SG_PACK
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
    SG_CHECK_OP16(dtype);
    if (n <= 0) return 0;
    int grid = (int)(cdiv(n, 256) < 148 * 16 ? cdiv(n, 256) : 148 * 16);
    if (is_op16(dtype))
        cast_f32_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(in, (__nv_bfloat16*)out, n);
    else
        cast_f32_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(in, (float*)out, n);
    return check_launch("cast_f32");
}

int sg_scale_f64_to_f32(const double* in, float* out, double scale, int n, void* stream) {
    scale_f64_to_f32_kernel<<<(n + 63) / 64, 64, 0, as_stream(stream)>>>(in, out, scale, n);
    return check_launch("scale_f64_to_f32");
}

int sg_gn_stats(const float* y, double* ws, float* mr, int C, int B, int T, int Tp, int G, void* stream) {
    SG_REQUIRE(G > 0 && C % G == 0, "gn_stats: C=%d not divisible by G=%d", C, G);
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
    int Cg = C / G;
    int rpb = Cg < 64 ? Cg : 64;
    dim3 grid((unsigned)cdiv(Cg, rpb), G, B);
    gn_stats_kernel<<<grid, kThreads, 0, st>>>(y, ws, C, B, T, Tp, G, rpb);
    gn_finalize_kernel<<<(B * G + 127) / 128, 128, 0, st>>>(ws, mr, B * G, 1.0 / ((double)Cg * T));
    return check_launch("gn_stats");
}

int sg_gn_act_fwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* res,
                  int res_is_f32, float res_scale, int act, int post_gelu, void* out_op, int planes,
                  long long plane_stride, float* out_f32, int C, int B, int T, int Tp, int G, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CHECK_OP16(y_dtype);
    SG_REQUIRE(Tp % 8 == 0, "gn_act_fwd: Tp %% 8 != 0");
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "gn_act_fwd: planes must be 1, 3 or 5");
    SG_REQUIRE(mr == nullptr || (G > 0 && C % G == 0), "gn_act_fwd: bad groups");
    SG_REQUIRE(y_dtype == SG_F32 || is_op16(dtype), "gn_act_fwd: a 16-bit y needs a 16-bit operand mode");
    cudaStream_t st = as_stream(stream);
    if (G <= 0) G = 1;
    typedef __nv_bfloat16 h16;
    const bool y16 = is_op16(y_dtype), r16 = res != nullptr && !res_is_f32;
    SG_REQUIRE(!r16 || is_op16(dtype), "gn_act_fwd: fp32 mode takes an fp32 residual");
    if (Tp == 8 && aligned16(y) && aligned16(res) && aligned16(out_op) && aligned16(out_f32) && (plane_stride & 7) == 0) {
        const int g = mr != nullptr ? G : 1;                 // short rows: one thread per row (static fields)
        gn_act_fwd_short_kernel<<<resident_grid(gn_act_fwd_short_kernel, short_task_blocks(C, B, g)), kThreads, 0, st>>>(y, y16, mr, gamma, beta, res, r16, res_scale, act,
                                                                              post_gelu, out_op, is_op16(dtype), planes,
                                                                              plane_stride, out_f32, C, B, T, G);
        return check_launch("gn_act_fwd");
    }
#define SG_FW(OT, RT, YT) return launch_fwd<OT, RT, YT>(act, post_gelu, y, mr, gamma, beta, res, res_scale, out_op, planes, plane_stride, out_f32, C, B, T, Tp, G, st)
    if (is_op16(dtype)) {
        if (y16) { if (r16) SG_FW(h16, h16, h16); else SG_FW(h16, float, h16); }
        else     { if (r16) SG_FW(h16, h16, float); else SG_FW(h16, float, float); }
    }
    SG_REQUIRE(!r16, "gn_act_fwd: fp32 mode takes an fp32 residual");
    SG_FW(float, float, float);
#undef SG_FW
}

int sg_gn_act_bwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* res,
                  int res_is_f32, float res_scale, int act, int post_gelu, void* dout, int dout_dtype, void* dy, int planes,
                  long long plane_stride, float* dgamma, float* dbeta, float* dbias, float* dres, int dres_accumulate,
                  double* ws, int C, int B, int T, int Tp, int G, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CHECK_OP16(y_dtype);
    SG_CHECK_OP16(dout_dtype);
    SG_REQUIRE(Tp % 8 == 0, "gn_act_bwd: Tp %% 8 != 0");
    SG_REQUIRE(planes == 1 || planes == 3 || planes == 5, "gn_act_bwd: planes must be 1, 3 or 5");
    SG_REQUIRE(mr == nullptr || (G > 0 && C % G == 0 && dgamma && dbeta && ws), "gn_act_bwd: bad GN arguments");
    SG_REQUIRE((y_dtype == SG_F32 && dout_dtype == SG_F32) || (is_op16(dtype) && mr != nullptr),
               "gn_act_bwd: 16-bit y / dout only for GroupNorm layers in a 16-bit operand mode");
    if (G <= 0) G = 1;
    cudaStream_t st = as_stream(stream);
    // dres_accumulate bit 1: the caller hands in dbias / dgamma / dbeta / ws already zeroed (one memset per step for
    // the whole gradient arena instead of four tiny ones per layer)
    const bool prezeroed = (dres_accumulate & 2) != 0;
    dres_accumulate &= 1;
    if (!prezeroed) {
        if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * C, st);
        if (mr != nullptr) {
            cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
            cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
            cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
        }
    }
    typedef __nv_bfloat16 h16;
    const bool r16 = res != nullptr && !res_is_f32;
    if (mr != nullptr) {
        Pass1Args p{};
        p.y = y; p.dout = dout; p.mr = mr; p.gamma = gamma; p.beta = beta; p.res = res; p.res_scale = res_scale;
        p.C = C; p.B = B; p.T = T; p.Tp = Tp; p.G = G;
        const bool y16 = is_op16(y_dtype), d16 = is_op16(dout_dtype);
        SG_REQUIRE(!r16 || is_op16(dtype), "gn_act_bwd: fp32 mode takes an fp32 residual");
        if (Tp == 8 && aligned16(y) && aligned16(res) && aligned16(dout) && aligned16(dy) && aligned16(dres) &&
            (plane_stride & 7) == 0) {                        // short rows: one thread per row (static fields)
            const long long blocks = short_task_blocks(C, B, G);
            gn_bwd_pass1_short_kernel<<<resident_grid(gn_bwd_pass1_short_kernel, blocks), kThreads, 0, st>>>(p, y16, d16, r16, act, post_gelu, dgamma, dbeta, ws, dres,
                                                                 dres_accumulate);
            const float inv_n = (float)(1.0 / ((double)(C / G) * T));
            gn_bwd_pass2_short_kernel<<<resident_grid(gn_bwd_pass2_short_kernel, blocks), kThreads, 0, st>>>(y, y16, dout, d16, mr, gamma, ws, dy, is_op16(dtype), planes,
                                                                 plane_stride, dbias, C, B, T, G, inv_n);
            return check_launch("gn_act_bwd");
        }
#define SG_BG(OT, RT, YT, DT) return launch_bwd_gn<OT, RT, YT, DT>(act, post_gelu, p, (OT*)dy, planes, plane_stride, dgamma, dbeta, dbias, dres, dres_accumulate, ws, st)
        if (is_op16(dtype)) {
            if (r16) {
                if (y16) { if (d16) SG_BG(h16, h16, h16, h16); else SG_BG(h16, h16, h16, float); }
                else     { if (d16) SG_BG(h16, h16, float, h16); else SG_BG(h16, h16, float, float); }
            } else {
                if (y16) { if (d16) SG_BG(h16, float, h16, h16); else SG_BG(h16, float, h16, float); }
                else     { if (d16) SG_BG(h16, float, float, h16); else SG_BG(h16, float, float, float); }
            }
        }
        SG_REQUIRE(!r16, "gn_act_bwd: fp32 mode takes an fp32 residual");
        SG_BG(float, float, float, float);
#undef SG_BG
    }
    BwdArgs p{};
    p.y = (const float*)y; p.mr = nullptr; p.gamma = gamma; p.beta = beta; p.res = res; p.res_scale = res_scale;
    p.dout = (const float*)dout; p.C = C; p.B = B; p.T = T; p.Tp = Tp; p.G = G; p.inv_n = 0.0;
    if (is_op16(dtype)) {
        if (!r16)
            return launch_bwd_plain<h16, float>(act, post_gelu, p, (h16*)dy, planes, plane_stride, dbias, dres, dres_accumulate, ws, st);
        return launch_bwd_plain<h16, h16>(act, post_gelu, p, (h16*)dy, planes, plane_stride, dbias, dres, dres_accumulate, ws, st);
    }
    return launch_bwd_plain<float, float>(act, post_gelu, p, (float*)dy, planes, plane_stride, dbias, dres, dres_accumulate, ws, st);
}

int sg_recon_fwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* xv,
                 int x_dtype, float* x_hat, double* loss_sums, float* rowsums, int N, int B, int T, int Tp, int G,
                 int loss_kind, void* stream) {
    SG_CHECK_OP16(y_dtype);
    SG_CHECK_OP16(x_dtype);
    SG_REQUIRE(G > 0 && N % G == 0, "recon_fwd: N=%d not divisible by G=%d", N, G);
    SG_REQUIRE(rowsums == nullptr || (xv != nullptr && aligned16(rowsums)), "recon_fwd: rowsums needs x and 16-byte alignment");
    const bool x16 = xv != nullptr && is_op16(x_dtype);
    SG_REQUIRE(!x16 || ((T & 7) == 0 && is_op16(y_dtype) && (long long)N * B < (1LL << 31)),
               "recon_fwd: a 16-bit packed target needs T %% 8 == 0 and a 16-bit y");
    const float* x = x16 ? nullptr : (const float*)xv;
    cudaStream_t st = as_stream(stream);
    if (xv != nullptr) cudaMemsetAsync(loss_sums, 0, sizeof(double) * 2, st);
    bool vec = (T % 4 == 0) && aligned16(x) && aligned16(x_hat);
    int grid = persistent_grid((long long)N * B) * 2;
    const bool mse = loss_kind == SG_LOSS_MSE;
    const bool ybf = is_op16(y_dtype);
    if (Tp == 8 && aligned16(y) && (!x16 || aligned16(xv))) {   // short rows: one thread per row (static fields)
#define SG_SH(YT, XT, MSE, RS, XH) \
    recon_fwd_short_kernel<YT, XT, MSE, RS, XH><<<resident_grid(recon_fwd_short_kernel<YT, XT, MSE, RS, XH>, short_tiles(N, B)), kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, (const XT*)xv, x_hat, loss_sums, \
                                                                           (float4*)rowsums, N, B, T, G, loss_kind)
#define SG_SH2(YT, XT, MSE) do { \
        if (rowsums) { if (x_hat) SG_SH(YT, XT, MSE, true, true); else SG_SH(YT, XT, MSE, true, false); } \
        else         { if (x_hat) SG_SH(YT, XT, MSE, false, true); else SG_SH(YT, XT, MSE, false, false); } } while (0)
        if (x16)      { if (mse) SG_SH2(__nv_bfloat16, __nv_bfloat16, true); else SG_SH2(__nv_bfloat16, __nv_bfloat16, false); }
        else if (ybf) { if (mse) SG_SH2(__nv_bfloat16, float, true); else SG_SH2(__nv_bfloat16, float, false); }
        else          { if (mse) SG_SH2(float, float, true); else SG_SH2(float, float, false); }
#undef SG_SH2
#undef SG_SH
        return check_launch("recon_fwd");
    }
    if (vec && xv != nullptr && (T & 7) == 0 && (long long)N * B < (1LL << 31)) {
        grid = persistent_grid(N);                           // one warp per channel
#define SG_FAST(YT, XT, MSE, RS, XH) \
    recon_fwd_fast_kernel<YT, XT, MSE, RS, XH><<<grid, kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, (const XT*)xv, x_hat, loss_sums, \
                                                                          (float4*)rowsums, N, B, T, Tp, G, loss_kind)
#define SG_FAST2(YT, XT, MSE) do { \
        if (rowsums) { if (x_hat) SG_FAST(YT, XT, MSE, true, true); else SG_FAST(YT, XT, MSE, true, false); } \
        else         { if (x_hat) SG_FAST(YT, XT, MSE, false, true); else SG_FAST(YT, XT, MSE, false, false); } } while (0)
        if (x16)      { if (mse) SG_FAST2(__nv_bfloat16, __nv_bfloat16, true); else SG_FAST2(__nv_bfloat16, __nv_bfloat16, false); }
        else if (ybf) { if (mse) SG_FAST2(__nv_bfloat16, float, true); else SG_FAST2(__nv_bfloat16, float, false); }
        else          { if (mse) SG_FAST2(float, float, true); else SG_FAST2(float, float, false); }
#undef SG_FAST2
#undef SG_FAST
        return check_launch("recon_fwd");
    }
    SG_REQUIRE(!x16, "recon_fwd: the 16-bit packed target is only supported by the fast path");
#define SG_RFWD(YT, VEC, MSE)                                                                                              \
    recon_fwd_kernel<YT, VEC, MSE><<<grid, kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, x, x_hat, loss_sums,          \
                                                              (float4*)rowsums, N, B, T, Tp, G, loss_kind)
#define SG_RFWD2(YT) do { \
        if (vec) { if (mse) SG_RFWD(YT, true, true); else SG_RFWD(YT, true, false); } \
        else     { if (mse) SG_RFWD(YT, false, true); else SG_RFWD(YT, false, false); } } while (0)
    if (ybf) SG_RFWD2(__nv_bfloat16); else SG_RFWD2(float);
#undef SG_RFWD2
#undef SG_RFWD
    return check_launch("recon_fwd");
}

int sg_recon_bwd(const void* y, int y_dtype, const float* mr, const float* gamma, const float* beta, const void* xv,
                 int x_dtype, const float* g_loss, const float* g_mse, float inv_numel, const float* dxhat_ext,
                 const float* rowsums, void* dy, float* dgamma, float* dbeta, float* dbias, double* ws, int N, int B, int T,
                 int Tp, int G, int loss_kind, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CHECK_OP16(x_dtype);
    const bool x16 = xv != nullptr && is_op16(x_dtype);
    SG_REQUIRE(!x16 || (rowsums != nullptr && dxhat_ext == nullptr && (T & 7) == 0 && is_op16(y_dtype) && is_op16(dtype)),
               "recon_bwd: a 16-bit packed target needs the one-pass path (row sums, T %% 8 == 0, 16-bit y and dy)");
    const float* x = x16 ? nullptr : (const float*)xv;
    SG_REQUIRE(G > 0 && N % G == 0 && Tp % 8 == 0, "recon_bwd: bad shape");
    SG_REQUIRE(xv != nullptr || (g_loss == nullptr && g_mse == nullptr), "recon_bwd: loss gradient without x");
    SG_REQUIRE(y_dtype == SG_F32 || y_dtype == dtype, "recon_bwd: a 16-bit y needs the same 16-bit mode");
    cudaStream_t st = as_stream(stream);
    double inv_n = 1.0 / ((double)(N / G) * T);
    // workspace: 2*B*G doubles for S followed by 2 floats for the folded scalars
    double* S = ws;
    float* scal = reinterpret_cast<float*>(ws + 2 * (size_t)B * G);
    cudaMemsetAsync(S, 0, sizeof(double) * 2 * B * G, st);
    cudaMemsetAsync(dbias, 0, sizeof(float) * N, st);
    recon_scalars_kernel<<<1, 1, 0, st>>>(g_loss, g_mse, inv_numel, scal);
    int grid = persistent_grid((long long)N * B) * 2;
    const bool ybf = is_op16(y_dtype);
    typedef __nv_bfloat16 bf;
    if (rowsums != nullptr && dxhat_ext == nullptr && xv != nullptr) {
        // one pass over y / x: the reductions of the GroupNorm backward were taken by the forward
        SG_REQUIRE((size_t)B * 2 * sizeof(float) <= 48 * 1024, "recon_bwd: batch too large for the combine kernel");
        dim3 gc((unsigned)cdiv(N / G, 64), G);
        recon_bwd_combine_kernel<<<gc, kThreads, sizeof(float) * 2 * B, st>>>((const float4*)rowsums, scal, gamma, dgamma,
                                                                               dbeta, S, N, B, G);
        bool vec = (T % 4 == 0) && aligned16(x);
        const bool mse = loss_kind == SG_LOSS_MSE;
        if (Tp == 8 && aligned16(y) && (!x16 || aligned16(xv)) && aligned16(dy)) {   // short rows: one thread per row
#define SG_AS(YT, XT, OT, MSE) \
    recon_bwd_apply_short_kernel<YT, XT, OT, MSE><<<resident_grid(recon_bwd_apply_short_kernel<YT, XT, OT, MSE>, short_tiles(N, B)), kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, (const XT*)xv, scal, S, (OT*)dy, dbias, \
                                                                             N, B, T, G, loss_kind, (float)inv_n)
            if (x16) {
                if (mse) SG_AS(bf, bf, bf, true); else SG_AS(bf, bf, bf, false);
            } else if (is_op16(dtype)) {
                if (ybf) { if (mse) SG_AS(bf, float, bf, true); else SG_AS(bf, float, bf, false); }
                else     { if (mse) SG_AS(float, float, bf, true); else SG_AS(float, float, bf, false); }
            } else {
                if (mse) SG_AS(float, float, float, true); else SG_AS(float, float, float, false);
            }
#undef SG_AS
            return check_launch("recon_bwd");
        }
        if (vec && (T & 7) == 0 && (long long)N * B < (1LL << 31)) {
            grid = persistent_grid(N);                       // one warp per channel; dbias written, not accumulated
#define SG_AF(YT, XT, OT, MSE) \
    recon_bwd_apply_fast_kernel<YT, XT, OT, MSE><<<grid, kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, (const XT*)xv, scal, S, (OT*)dy, dbias, \
                                                                            N, B, T, Tp, G, loss_kind, (float)inv_n)
            if (x16) {
                if (mse) SG_AF(bf, bf, bf, true); else SG_AF(bf, bf, bf, false);
            } else if (is_op16(dtype)) {
                if (ybf) { if (mse) SG_AF(bf, float, bf, true); else SG_AF(bf, float, bf, false); }
                else     { if (mse) SG_AF(float, float, bf, true); else SG_AF(float, float, bf, false); }
            } else {
                if (mse) SG_AF(float, float, float, true); else SG_AF(float, float, float, false);
            }
#undef SG_AF
            return check_launch("recon_bwd");
        }
        SG_REQUIRE(!x16, "recon_bwd: the 16-bit packed target is only supported by the fast path");
#define SG_APPLY(YT, OT, VEC, MSE)                                                                                         \
    recon_bwd_apply_kernel<YT, OT, VEC, MSE><<<grid, kThreads, 0, st>>>((const YT*)y, mr, gamma, beta, x, scal, S, (OT*)dy,  \
                                                                         dbias, N, B, T, Tp, G, loss_kind, inv_n)
#define SG_APPLY2(YT, OT) do { \
        if (vec) { if (mse) SG_APPLY(YT, OT, true, true); else SG_APPLY(YT, OT, true, false); } \
        else     { if (mse) SG_APPLY(YT, OT, false, true); else SG_APPLY(YT, OT, false, false); } } while (0)
        if (is_op16(dtype)) { if (ybf) SG_APPLY2(bf, bf); else SG_APPLY2(float, bf); }
        else                  SG_APPLY2(float, float);
#undef SG_APPLY2
#undef SG_APPLY
        return check_launch("recon_bwd");
    }
    ReconBwdArgs p{};
    p.y = y; p.mr = mr; p.gamma = gamma; p.beta = beta; p.x = x; p.ext = dxhat_ext; p.loss_kind = loss_kind;
    p.N = N; p.B = B; p.T = T; p.Tp = Tp; p.G = G; p.inv_n = inv_n;
    cudaMemsetAsync(dgamma, 0, sizeof(float) * N, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * N, st);
#define SG_TWO(YT, OT)                                                                                             \
    do {                                                                                                           \
        recon_bwd_kernel<YT, OT, true><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (OT*)dy, dbias);       \
        recon_bwd_kernel<YT, OT, false><<<grid, kThreads, 0, st>>>(p, scal, dgamma, dbeta, S, (OT*)dy, dbias);      \
    } while (0)
    if (is_op16(dtype)) { if (ybf) SG_TWO(bf, bf); else SG_TWO(float, bf); }
    else                  SG_TWO(float, float);
#undef SG_TWO
    return check_launch("recon_bwd");
}

}  // extern "C"
