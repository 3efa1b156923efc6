// Batched spectral-norm preparation: power iteration + sigma + normalised bf16/fp32 GEMM-layout weights
// for EVERY layer of a sub-network in five launches (common.py:15-37 -> spectral_norm.py:62-114).
//
// The power iteration depends only on the weights and the stored u / v, never on activations, so the
// ~40 layers of the encoder / decoder are prepared up front instead of layer by layer inside the forward
// (which costs ~8 launches and two under-filled weight streams per layer).  HBM traffic: W is read three
// times (W^T u, W v, pack) and the operand copy written once - 14 B per weight element.
//
//   phase 1 (training)  vraw = W^T u                       tiles of 64 rows x 1024 columns, atomics on vraw
//   phase 2             uraw = W vin  (vin = vraw | v)     one warp per (row, 4096-column chunk)
//   phase 3             v = vraw/|vraw|, u = normalize(uraw/|vraw|), sigma = u . (W v)      one CTA per layer
//   phase 4             Wg[j'][o][i] = W(o,i,j) / sigma    one thread per (o, 8 consecutive i, all taps)
//
// Normalisation is scale-invariant, so phase 2 can run on the un-normalised vraw: W v = uraw / |vraw|.
#include "common.cuh"

namespace sg {

constexpr int kMaxLayers = 96;
constexpr int kP1Rows = 64, kP1Cols = 1024;
constexpr int kP2Cols = 4096;
constexpr int kSnThreads = 256;

struct SnPrefix {
    int n;
    int start[kMaxLayers + 1];
};

__device__ __forceinline__ int sn_find(const SnPrefix& pf, int chunk) {
    int lo = 0, hi = pf.n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pf.start[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool row_contiguous(const sg_sn_layer& L) { return L.so == (long long)L.Cin * L.k && L.si == L.k; }
__device__ __forceinline__ long long w_addr(const sg_sn_layer& L, int o, int q) {
    int i = q / L.k, j = q - i * L.k;
    return (long long)o * L.so + (long long)i * L.si + j;
}

// ---- phase 1: vraw[q] += sum_{o in row block} W[o][q] u[o] -------------------------------------------
__global__ void __launch_bounds__(kSnThreads)
sn_p1_kernel(const sg_sn_layer* __restrict__ layers, const __grid_constant__ SnPrefix pf) {
    const int li = sn_find(pf, blockIdx.x);
    const sg_sn_layer L = layers[li];
    const int Wd = L.Cin * L.k;
    const int cblocks = (Wd + kP1Cols - 1) / kP1Cols;
    const int local = blockIdx.x - pf.start[li];
    const int rb = local / cblocks, cb = local - rb * cblocks;
    const int o_lo = rb * kP1Rows, o_hi = min(L.H, o_lo + kP1Rows);
    const int q0 = cb * kP1Cols + threadIdx.x * 4;
    if (q0 >= Wd) return;
    float* vraw = L.ws;
    if (row_contiguous(L) && (Wd & 3) == 0 && ((reinterpret_cast<uintptr_t>(L.w) & 15) == 0)) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = L.w + q0;
#pragma unroll 8
        for (int o = o_lo; o < o_hi; ++o) {
            float4 w = __ldg(reinterpret_cast<const float4*>(base + (long long)o * Wd));
            float uo = __ldg(L.u + o);
            acc.x += w.x * uo; acc.y += w.y * uo; acc.z += w.z * uo; acc.w += w.w * uo;
        }
        atomicAdd(vraw + q0, acc.x);
        atomicAdd(vraw + q0 + 1, acc.y);
        atomicAdd(vraw + q0 + 2, acc.z);
        atomicAdd(vraw + q0 + 3, acc.w);
    } else {
        for (int d = 0; d < 4; ++d) {
            int q = q0 + d;
            if (q >= Wd) break;
            float acc = 0.f;
            for (int o = o_lo; o < o_hi; ++o) acc += __ldg(L.w + w_addr(L, o, q)) * __ldg(L.u + o);
            atomicAdd(vraw + q, acc);
        }
    }
}

// ---- phase 2: uraw[o] += sum_{q in chunk} W[o][q] vin[q] ----------------------------------------------
__global__ void __launch_bounds__(kSnThreads)
sn_p2_kernel(const sg_sn_layer* __restrict__ layers, const __grid_constant__ SnPrefix pf, int training) {
    const int li = sn_find(pf, blockIdx.x);
    const sg_sn_layer L = layers[li];
    const int Wd = L.Cin * L.k;
    const int cchunks = (Wd + kP2Cols - 1) / kP2Cols;
    const int local = blockIdx.x - pf.start[li];
    const int rg = local / cchunks, cc = local - rg * cchunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = rg * 8 + warp;
    if (o >= L.H) return;
    const float* vin = training ? L.ws : L.v;
    float* uraw = L.ws + Wd;
    const int q_lo = cc * kP2Cols, q_hi = min(Wd, q_lo + kP2Cols);
    float acc = 0.f;
    if (row_contiguous(L) && (Wd & 3) == 0 && ((reinterpret_cast<uintptr_t>(L.w) & 15) == 0) &&
        ((reinterpret_cast<uintptr_t>(vin) & 15) == 0)) {
        const float* row = L.w + (long long)o * Wd;
        for (int q = q_lo + lane * 4; q < q_hi; q += 128) {
            float4 w = __ldg(reinterpret_cast<const float4*>(row + q));
            float4 x = *reinterpret_cast<const float4*>(vin + q);
            acc += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
        }
    } else {
        for (int q = q_lo + lane; q < q_hi; q += 32) acc += __ldg(L.w + w_addr(L, o, q)) * vin[q];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        if (cchunks == 1) uraw[o] = acc; else atomicAdd(uraw + o, acc);
    }
}

// ---- phase 3: one CTA per layer ------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) sn_p3_kernel(const sg_sn_layer* __restrict__ layers, int training) {
    __shared__ double sh[32];
    __shared__ float s_a, s_b;
    const sg_sn_layer L = layers[blockIdx.x];
    if (!L.has_sn) {
        if (threadIdx.x == 0) L.sigma[0] = 1.f;
        return;
    }
    const int Wd = L.Cin * L.k;
    const float* vraw = L.ws;
    const float* uraw = L.ws + Wd;
    if (training) {
        double s = 0.0;
        for (int q = threadIdx.x; q < Wd; q += blockDim.x) s += (double)vraw[q] * (double)vraw[q];
        double t = block_sum(s, sh);
        if (threadIdx.x == 0) s_a = 1.f / fmaxf((float)sqrt(t), 1e-12f);
        __syncthreads();
        const float inv_v = s_a;
        for (int q = threadIdx.x; q < Wd; q += blockDim.x) L.v[q] = vraw[q] * inv_v;
        s = 0.0;
        for (int o = threadIdx.x; o < L.H; o += blockDim.x) {
            double wv = (double)(uraw[o] * inv_v);
            s += wv * wv;
        }
        t = block_sum(s, sh);
        if (threadIdx.x == 0) {
            float nrm = fmaxf((float)sqrt(t), 1e-12f);
            s_b = 1.f / nrm;
            L.sigma[0] = (float)(t / (double)nrm);
        }
        __syncthreads();
        const float inv_u = s_b;
        for (int o = threadIdx.x; o < L.H; o += blockDim.x) L.u[o] = uraw[o] * inv_v * inv_u;
    } else {
        double s = 0.0;
        for (int o = threadIdx.x; o < L.H; o += blockDim.x) s += (double)uraw[o] * (double)L.u[o];
        double t = block_sum(s, sh);
        if (threadIdx.x == 0) L.sigma[0] = (float)t;
    }
}

// ---- phase 4: operand copy ---------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(kSnThreads)
sn_p4_kernel(const sg_sn_layer* __restrict__ layers, const __grid_constant__ SnPrefix pf) {
    const int li = sn_find(pf, blockIdx.x);
    const sg_sn_layer L = layers[li];
    const int octs = L.Cin_p >> 3;
    const long long t = (long long)(blockIdx.x - pf.start[li]) * kSnThreads + threadIdx.x;
    if (t >= (long long)L.H * octs) return;
    const int o = (int)(t / octs), i0 = (int)(t - (long long)o * octs) * 8;
    const float inv = 1.f / L.sigma[0];
    OT* wg = reinterpret_cast<OT*>(L.wg);
    const int k = L.k;
    if (k == 1 && L.si == 1 && i0 + 8 <= L.Cin && (((reinterpret_cast<uintptr_t>(L.w) >> 2) + (long long)o * L.so + i0) & 3) == 0) {
        F8 r = load8(L.w + (long long)o * L.so + i0);
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] *= inv;
        store8(wg + (long long)o * L.Cin_p + i0, r);
        return;
    }
    for (int j = 0; j < k; ++j) {
        F8 r;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int i = i0 + e;
            r.v[e] = i < L.Cin ? __ldg(L.w + (long long)o * L.so + (long long)i * L.si + j) * inv : 0.f;
        }
        int jj = L.flip ? k - 1 - j : j;
        store8(wg + ((long long)jj * L.H + o) * L.Cin_p + i0, r);
    }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_sn_prepare(const sg_sn_layer* layers_dev, const sg_sn_layer* layers_host, int n_layers, float* ws_base,
                             long long ws_elems, int training, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_REQUIRE(n_layers > 0 && n_layers <= kMaxLayers, "sn_prepare: n_layers=%d out of range (max %d)", n_layers, kMaxLayers);
    cudaStream_t st = as_stream(stream);
    SnPrefix p1, p2, p4;
    p1.n = p2.n = p4.n = n_layers;
    long long t1 = 0, t2 = 0, t4 = 0;
    bool any_sn = false;
    for (int i = 0; i < n_layers; ++i) {
        const sg_sn_layer& L = layers_host[i];
        long long Wd = (long long)L.Cin * L.k;
        SG_REQUIRE(L.H > 0 && Wd > 0 && Wd < (1LL << 31), "sn_prepare: layer %d has a bad shape", i);
        SG_REQUIRE(!L.has_wg || (L.Cin_p % 8 == 0 && L.Cin_p >= L.Cin), "sn_prepare: layer %d: bad Cin_p", i);
        p1.start[i] = (int)t1;
        p2.start[i] = (int)t2;
        p4.start[i] = (int)t4;
        if (L.has_sn) {
            any_sn = true;
            SG_REQUIRE(L.ws != nullptr && L.u != nullptr && L.v != nullptr, "sn_prepare: layer %d misses u/v/ws", i);
            t1 += cdiv(L.H, kP1Rows) * cdiv(Wd, kP1Cols);
            t2 += cdiv(L.H, 8) * cdiv(Wd, kP2Cols);
        }
        if (L.has_wg) t4 += cdiv((long long)L.H * (L.Cin_p / 8), kSnThreads);
    }
    p1.start[n_layers] = (int)t1;
    p2.start[n_layers] = (int)t2;
    p4.start[n_layers] = (int)t4;
    SG_REQUIRE(t1 < (1LL << 31) && t2 < (1LL << 31) && t4 < (1LL << 31), "sn_prepare: grid too large");
    if (any_sn) {
        cudaMemsetAsync(ws_base, 0, sizeof(float) * (size_t)ws_elems, st);
        if (training && t1 > 0) sn_p1_kernel<<<(unsigned)t1, kSnThreads, 0, st>>>(layers_dev, p1);
        if (t2 > 0) sn_p2_kernel<<<(unsigned)t2, kSnThreads, 0, st>>>(layers_dev, p2, training);
    }
    sn_p3_kernel<<<n_layers, 1024, 0, st>>>(layers_dev, training);
    if (t4 > 0) {
        if (is_op16(dtype)) sn_p4_kernel<__nv_bfloat16><<<(unsigned)t4, kSnThreads, 0, st>>>(layers_dev, p4);
        else sn_p4_kernel<float><<<(unsigned)t4, kSnThreads, 0, st>>>(layers_dev, p4);
    }
    return check_launch("sn_prepare");
}
