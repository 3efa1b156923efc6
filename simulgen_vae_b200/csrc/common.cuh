// Shared device helpers for the simulgen_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/simulgen_b200.h"

namespace sg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SG_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            sg::set_error(__VA_ARGS__);       \
            return 1;                         \
        }                                     \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool is_op16(int dtype) { return dtype == SG_BF16 || dtype == SG_F16; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

constexpr float kGnEps = 1e-5f;


// ---------------------------------------------------------------------------------------------
// 8-element vector access for fp32 / bf16 rows (rows are 8-element aligned: Tp % 8 == 0)
// ---------------------------------------------------------------------------------------------
struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 load8(const float* p) {
    F8 r;
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
// 16-bit operands.  Every 16-bit tensor of the engine is typed __nv_bfloat16 in the kernels; whether the bits are
// bf16 or IEEE fp16 ("fp16" precision mode: 3 more mantissa bits, tcgen05 kind::f16 runs both at the same rate) is a
// compile-time constant of the LIBRARY: the sources are built twice (libsimulgen_b200.so and, with -DSG_OP16_HALF,
// libsimulgen_b200_fp16.so) and the Python side calls the one that matches the operand dtype.  (A run-time switch in
// constant memory was measured at -1.6 % step throughput: these kernels are issue-bound.)
#ifdef SG_OP16_HALF
constexpr bool c_op16_half = true;
#else
constexpr bool c_op16_half = false;
#endif
// every entry point that takes a dtype checks that this library variant handles it
#define SG_CHECK_OP16(dtype)                                                                                         \
    SG_REQUIRE((dtype) != (sg::c_op16_half ? SG_BF16 : SG_F16), "this build of libsimulgen_b200 handles %s operands", \
               sg::c_op16_half ? "fp16 (and fp32)" : "bf16 (and fp32)")

__device__ __forceinline__ float2 op16x2_to_f2(uint32_t bits) {
    if (c_op16_half) return __half22float2(*reinterpret_cast<const __half2*>(&bits));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bits));
}
__device__ __forceinline__ uint32_t f2_to_op16x2(float a, float b) {
    if (c_op16_half) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
    F8 r;
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = op16x2_to_f2(w[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
// Raw (unconverted) 8-element loads: kernels that keep several rows in flight hold the 16-bit rows as 4 registers each
// until the row is actually processed (holding them converted would double the register footprint of the prefetch).
struct Raw16 { uint4 q; };
struct Raw32 { float4 a, b; };
__device__ __forceinline__ Raw16 load_raw(const __nv_bfloat16* p) {
    Raw16 r;
    r.q = *reinterpret_cast<const uint4*>(p);
    return r;
}
__device__ __forceinline__ Raw32 load_raw(const float* p) {
    Raw32 r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
}
__device__ __forceinline__ F8 cvt8(const Raw16& w) {
    F8 r;
    const uint32_t* u = reinterpret_cast<const uint32_t*>(&w.q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = op16x2_to_f2(u[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ F8 cvt8(const Raw32& w) {
    F8 r;
    r.v[0] = w.a.x; r.v[1] = w.a.y; r.v[2] = w.a.z; r.v[3] = w.a.w;
    r.v[4] = w.b.x; r.v[5] = w.b.y; r.v[6] = w.b.z; r.v[7] = w.b.w;
    return r;
}
template <typename T> struct RawOf;
template <> struct RawOf<float> { typedef Raw32 type; };
template <> struct RawOf<__nv_bfloat16> { typedef Raw16 type; };

__device__ __forceinline__ void store8(float* p, const F8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
    uint4 raw;
    uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = f2_to_op16x2(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) {
    if (c_op16_half) return __half2float(*reinterpret_cast<const __half*>(&x));
    return __bfloat162float(x);
}
__device__ __forceinline__ void from_f(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float x) {
    if (c_op16_half) {
        __half h = __float2half_rn(x);
        d = *reinterpret_cast<__nv_bfloat16*>(&h);
    } else {
        d = __float2bfloat16_rn(x);
    }
}

// ---------------------------------------------------------------------------------------------
// shifted operand planes.  TMA moves 16-byte granules, so a conv tap cannot be a 1-element shift of
// the box along the contiguous r axis.  Operands that feed a k-tap conv are therefore stored as k
// "planes": plane[pl][c][b][t] = value[c][b][t + pl - planes/2] (zero outside [0, T)), and tap j of
// the conv simply selects a plane (an aligned TMA coordinate).  Rows are independent: the shift
// never crosses a sample because the gap t >= T is zero.
//
// srow: per-warp smem buffer of Tp + 8 floats; the row lives at srow[4 + t] (zero for t >= T) with
// 4 zero floats of halo on both sides.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void srow_clear_halo(float* srow, int Tp, int lane) {
    if (lane < 4) srow[lane] = 0.f;
    else if (lane < 8) srow[Tp + lane] = 0.f;
}

template <typename OT>
__device__ __forceinline__ void store_row_planes(OT* base, long long row_off, int planes, long long pstride,
                                                 const float* srow, int T, int Tp, int lane) {
    const int half = planes / 2;
    for (int pl = 0; pl < planes; ++pl) {
        const int s = pl - half;
        OT* dst = base + (long long)pl * pstride + row_off;
        for (int seg = lane; seg < Tp / 8; seg += 32) {
            F8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int t = seg * 8 + i;
                o.v[i] = (t < T) ? srow[4 + t + s] : 0.f;
            }
            store8(dst + seg * 8, o);
        }
    }
}

// Register / shuffle variant of store_row_planes for rows of at most 256 elements (one 8-element segment per
// lane): the two neighbours a shifted plane needs on each side come from the adjacent lanes by warp shuffle, so
// the k shifted copies cost 4 shuffles instead of a shared-memory round trip per element and plane (the smem
// version made the 5-plane writers MIO-bound).  Must be called by all 32 lanes; `o` holds this lane's segment
// (zeros for t >= T and for lanes beyond the row), lanes with seg >= nseg_p do not store.
template <typename OT, int PLANES>
__device__ __forceinline__ void store_planes_shfl(OT* base, long long row_off, long long pstride, const F8& o, int T, int nseg_p,
                                                  int lane) {
    float ext[12];
    ext[0] = __shfl_up_sync(0xffffffffu, o.v[6], 1);
    ext[1] = __shfl_up_sync(0xffffffffu, o.v[7], 1);
    ext[10] = __shfl_down_sync(0xffffffffu, o.v[0], 1);
    ext[11] = __shfl_down_sync(0xffffffffu, o.v[1], 1);
    if (lane == 0) { ext[0] = 0.f; ext[1] = 0.f; }
    if (lane == 31) { ext[10] = 0.f; ext[11] = 0.f; }
#pragma unroll
    for (int i = 0; i < 8; ++i) ext[2 + i] = o.v[i];
    if (lane >= nseg_p) return;
    const int t0 = lane * 8;
#pragma unroll
    for (int pl = 0; pl < PLANES; ++pl) {
        const int s = pl - PLANES / 2;
        F8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = (t0 + i < T) ? ext[2 + i + s] : 0.f;
        store8(base + (long long)pl * pstride + row_off + t0, r);
    }
}
template <typename OT>
__device__ __forceinline__ void store_planes_shfl_n(OT* base, long long row_off, int planes, long long pstride, const F8& o, int T,
                                                    int nseg_p, int lane) {
    if (planes == 5) store_planes_shfl<OT, 5>(base, row_off, pstride, o, T, nseg_p, lane);
    else if (planes == 3) store_planes_shfl<OT, 3>(base, row_off, pstride, o, T, nseg_p, lane);
    else store_planes_shfl<OT, 1>(base, row_off, pstride, o, T, nseg_p, lane);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: 4 standard normals keyed on (seed, stream id, sample id, quad index)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t stream_id, uint64_t sample, uint64_t qd, float (&n)[4]) {
    uint32_t c0 = (uint32_t)qd, c1 = (uint32_t)(qd >> 32) ^ (uint32_t)(stream_id << 8);
    uint32_t c2 = (uint32_t)sample, c3 = (uint32_t)(sample >> 32) ^ (uint32_t)(stream_id >> 24);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const float k2_32 = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)c0 + 0.5f) * k2_32, u1 = ((float)c1 + 0.5f) * k2_32;
    float u2 = ((float)c2 + 0.5f) * k2_32, u3 = ((float)c3 + 0.5f) * k2_32;
    u0 = fminf(fmaxf(u0, 1e-10f), 1.0f);
    u2 = fminf(fmaxf(u2, 1e-10f), 1.0f);
    // MUFU-based log / sin / cos (abs error ~1e-6: far below the sampling noise of a Gaussian draw)
    float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
    float s0, co0, s1, co1;
    __sincosf(6.283185307179586f * u1, &s0, &co0);
    __sincosf(6.283185307179586f * u3, &s1, &co1);
    n[0] = r0 * co0; n[1] = r0 * s0; n[2] = r1 * co1; n[3] = r1 * s1;
}

// ---------------------------------------------------------------------------------------------
// mbarrier helpers (used by the tcgen05 GEMM pipelines)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("simulgen_b200: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// ---------------------------------------------------------------------------------------------
// reconstruction losses (VAE_network.py:71-77): per-element term and derivative wrt d = x_hat - x
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

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of a double; result valid in thread 0.  `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// activations (exact erf GELU = nn.GELU() default; nn.Tanh)
//
// erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32 round-off level): one MUFU.EX2, one
// MUFU.RCP and 6 FMAs, and the exponential exp(-x^2/2) is the very one the Gaussian pdf of GELU' needs, so
// value and derivative cost ~20 issue slots together.  erff() + expf() cost ~100 and made the GroupNorm /
// activation kernels ALU-bound instead of HBM-bound.
// ---------------------------------------------------------------------------------------------
// single-instruction MUFU approximations (the CUDA intrinsics __expf / __fdividef add denormal and range
// handling - FSETP / FSEL / extra FMULs per call - that these bounded arguments never need)
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
    const float e = fast_ex2(x * x * -0.72134752044448170368f);      // exp(-x^2 / 2)
    // 0.5 * (a1 t + ... + a5 t^5)
    float poly = fmaf(t, 0.5307027145f, -0.7265760135f);
    poly = fmaf(t, poly, 0.7107068705f);
    poly = fmaf(t, poly, -0.142248368f);
    poly = fmaf(t, poly, 0.127414796f);
    const float half_erfc = poly * t * e;                           // 0.5 * (1 - erf(|x| / sqrt 2))
    const float cdf = 0.5f + copysignf(0.5f - half_erfc, x);
    g = x * cdf;
    dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float gelu_f(float x) {
    float g, dg;
    gelu_both(x, g, dg);
    return g;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    float g, dg;
    gelu_both(x, g, dg);
    return dg;
}
// tanh(x) = 1 - 2 / (exp(2x) + 1): MUFU.EX2 + MUFU.RCP, abs error < 4e-7
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = fast_ex2(x * 2.88539008177792681472f);           // exp(2x); +inf for large x -> rcp = 0 -> 1
    return fmaf(-2.f, fast_rcp(e + 1.f), 1.f);
}
__device__ __forceinline__ float act_f(int act, float x) {
    return act == SG_ACT_GELU ? gelu_f(x) : (act == SG_ACT_TANH ? tanh_fast(x) : x);
}
// value and derivative of act at pre-activation x
__device__ __forceinline__ void act_both(int act, float x, float& val, float& dval) {
    if (act == SG_ACT_GELU) {
        gelu_both(x, val, dval);
    } else if (act == SG_ACT_TANH) {
        val = tanh_fast(x);
        dval = 1.0f - val * val;
    } else {
        val = x;
        dval = 1.0f;
    }
}

struct GnStat {
    float mean, rstd;
};
__device__ __forceinline__ GnStat gn_stat(const double* stats, int b, int g, int G, double inv_n) {
    double s = stats[(size_t)(b * G + g) * 2], ss = stats[(size_t)(b * G + g) * 2 + 1];
    double m = s * inv_n;
    double var = ss * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    GnStat r;
    r.mean = (float)m;
    r.rstd = (float)(1.0 / sqrt(var + (double)kGnEps));
    return r;
}

}  // namespace sg
