// Shared device helpers for the simulgen_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
    F8 r;
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ void store8(float* p, const F8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float x) { d = __float2bfloat16_rn(x); }

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
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
    float z = fabsf(x) * 0.70710678118654752440f;
    float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float e = __expf(-z * z);                                    // exp(-x^2 / 2)
    float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
    float half_erfc = 0.5f * poly * e;                           // 0.5 * (1 - erf(|x| / sqrt 2))
    float cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
    g = x * cdf;
    dg = cdf + x * 0.39894228040143267794f * e;
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
    float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
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
