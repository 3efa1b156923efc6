// Multi-tensor optimiser step: spectral-norm weight gradient + AdamW + global gradient norm in two
// launches for the whole model (train.py:92,156-168; torch/nn/utils/spectral_norm.py:97-113 backward).
//
// The per-layer path (sg_sn_weight_grad + sg_adamw_step) costs two passes over every weight gradient,
// a materialised gradient in the reference layout and ~250 launches per step.  Here the wgrad GEMM
// output G (fp32, GEMM layout [k][Cout][Cin_p]) is consumed directly:
//     dot_l  = <G_l, W_l>                                          (launch 1, all layers)
//     grad   = (G - (dot/sigma) u v^T) / sigma  ->  AdamW update   (launch 2, all tensors)
// Launch 2 walks every tensor in its NATIVE layout (p, m, v coalesced); G is gathered from its k tap
// planes.  HBM traffic per element: launch 1 reads G, W (8 B); launch 2 reads G, W, m, v and writes
// W, m, v (28 B) - the AdamW minimum.
#include "common.cuh"

namespace sg {

constexpr int kOptChunk = 8192;      // elements per CTA
constexpr int kOptThreads = 256;
constexpr int kMaxItems = 512;

struct OptPrefix {
    int n_items;
    int start[kMaxItems + 1];        // first chunk of item i; start[n_items] = total chunks
};

__device__ __forceinline__ int find_item(const OptPrefix& pf, int chunk) {
    int lo = 0, hi = pf.n_items;     // invariant: start[lo] <= chunk < start[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pf.start[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

// native flat index e of weight_orig -> (o, q = i*k + j, index into the GEMM-layout gradient)
__device__ __forceinline__ void decode(const sg_opt_item& it, int e, int& o, int& q, long long& gidx) {
    int i, j;
    if (it.k == 1 && !it.flip) {
        o = e / it.Cin;
        i = e - o * it.Cin;
        j = 0;
    } else if (!it.flip) {           // Conv1d [Cout][Cin][k]
        int wd = it.Cin * it.k;
        o = e / wd;
        int r = e - o * wd;
        i = r / it.k;
        j = r - i * it.k;
    } else {                         // ConvTranspose1d [Cin][Cout][k]
        int wd = it.Cout * it.k;
        i = e / wd;
        int r = e - i * wd;
        o = r / it.k;
        j = r - o * it.k;
    }
    q = i * it.k + j;
    int jj = it.flip ? it.k - 1 - j : j;
    gidx = ((long long)jj * it.Cout + o) * it.Cin_p + i;
}

// `bad` (may be NULL): incremented when a block sees a non-finite gradient.  For spectral-norm layers the block's share
// of <G, W> is non-finite whenever one of its G elements is (inf * w = +-inf or NaN, NaN * w = NaN, and sums keep it);
// plain vectors (biases, GroupNorm affine) are scanned with 0 * g, which is NaN exactly for non-finite g.
__global__ void __launch_bounds__(kOptThreads)
opt_dot_kernel(const sg_opt_item* __restrict__ items, const __grid_constant__ OptPrefix pf, double* __restrict__ bad) {
    __shared__ double sh[32];
    const int idx = find_item(pf, blockIdx.x);
    const sg_opt_item it = items[idx];
    const long long e0 = (long long)(blockIdx.x - pf.start[idx]) * kOptChunk;
    const long long e1 = min(it.n, e0 + kOptChunk);
    if (it.u == nullptr) {
        if (bad == nullptr) return;
        float z = 0.f;
        for (long long e = e0 + threadIdx.x; e < e1; e += kOptThreads) z = fmaf(0.f, it.g[e], z);
        double t = block_sum((double)z, sh);
        if (threadIdx.x == 0 && !isfinite(t)) atomicAdd(bad, 1.0);
        return;
    }
    float acc = 0.f;
    const bool direct = it.k == 1 && !it.flip && it.Cin_p == it.Cin;
    if (direct && (it.n & 3) == 0 && aligned16(it.g, it.p)) {
        for (long long e = e0 + threadIdx.x * 4; e < e1; e += kOptThreads * 4) {
            float4 g = *reinterpret_cast<const float4*>(it.g + e);
            float4 w = *reinterpret_cast<const float4*>(it.p + e);
            acc += g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
        }
    } else if (!it.flip && (it.n & 3) == 0 && aligned16(it.p, it.p)) {
        // k-tap Conv1d: p is walked 4 native elements per thread (16-byte loads), G is gathered from its tap planes
        for (long long e = e0 + threadIdx.x * 4; e < e1; e += kOptThreads * 4) {
            float4 w4 = *reinterpret_cast<const float4*>(it.p + e);
            float w[4] = {w4.x, w4.y, w4.z, w4.w};
            int o, q;
            long long gi;
            decode(it, (int)e, o, q, gi);
            int i = q / it.k, j = q - i * it.k;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                acc += __ldg(it.g + ((long long)j * it.Cout + o) * it.Cin_p + i) * w[t];
                if (++j == it.k) { j = 0; if (++i == it.Cin) { i = 0; ++o; } }
            }
        }
    } else {
        for (long long e = e0 + threadIdx.x; e < e1; e += kOptThreads) {
            int o, q;
            long long gi;
            decode(it, (int)e, o, q, gi);
            acc += it.g[gi] * it.p[e];
        }
    }
    double t = block_sum((double)acc, sh);
    if (threadIdx.x == 0) {
        atomicAdd(it.dot, t);
        if (bad != nullptr && !isfinite(t)) atomicAdd(bad, 1.0);
    }
}

// Peer memory (include/simulgen_b200.h, sg_peer): gradient of element `g` (a pointer into MY arena) on rank r
__device__ __forceinline__ const float* peer_grad(const sg_peer& pc, const float* g, bool vec_arena, int r) {
    const char* mine = vec_arena ? pc.vbase[pc.rank] : pc.wbase[pc.rank];
    const char* theirs = vec_arena ? pc.vbase[r] : pc.wbase[r];
    return reinterpret_cast<const float*>(theirs + (reinterpret_cast<const char*>(g) - mine));
}
__device__ __forceinline__ float* peer_param(const sg_peer& pc, float* p, int r) {
    return reinterpret_cast<float*>(pc.pbase[r] + (reinterpret_cast<char*>(p) - pc.pbase[pc.rank]));
}
// NVSwitch multicast (NVLS) addresses of the element `g` / `p` of MY buffers
__device__ __forceinline__ const float* mc_grad(const sg_peer& pc, const float* g, bool vec_arena) {
    const char* mine = vec_arena ? pc.vbase[pc.rank] : pc.wbase[pc.rank];
    return reinterpret_cast<const float*>((vec_arena ? pc.vmc : pc.wmc) + (reinterpret_cast<const char*>(g) - mine));
}
__device__ __forceinline__ float* mc_param(const sg_peer& pc, float* p) {
    return reinterpret_cast<float*>(pc.pmc + (reinterpret_cast<char*>(p) - pc.pbase[pc.rank]));
}
// sum over every rank's copy, reduced inside the switch
__device__ __forceinline__ float4 multimem_sum4(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ float multimem_sum1(const float* mc) {
    float r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(r) : "l"(mc) : "memory");
    return r;
}
// one store that lands in every rank's copy
__device__ __forceinline__ void multimem_st4(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st1(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}

// Fused reduce-scatter + dot over NVLink peer memory.  `items` describe this rank's shard of every tensor; for each of its
// elements the gradient is loaded from all `world` arenas (world - 1 of them remote: coalesced 16-byte P2P loads on the
// k = 1 layers, tap-plane gathers otherwise), summed, written back into this rank's arena - no other rank reads or writes
// those addresses: shards are disjoint in both layouts - and multiplied into the layer's <G, W>.
__global__ void __launch_bounds__(kOptThreads)
peer_reduce_dot_kernel(const sg_opt_item* __restrict__ items, const __grid_constant__ OptPrefix pf,
                       const __grid_constant__ sg_peer pc, double* __restrict__ bad) {
    __shared__ double sh[32];
    // grid-stride over the 8192-element chunks: the grid may be much smaller than the chunk count (a launch that runs
    // UNDERNEATH the backward pass keeps 2 blocks per SM so that the persistent GEMM CTAs still find room)
    for (int chunk = blockIdx.x; chunk < pf.start[pf.n_items]; chunk += gridDim.x) {
    const int idx = find_item(pf, chunk);
    const sg_opt_item it = items[idx];
    const long long e0 = (long long)(chunk - pf.start[idx]) * kOptChunk;
    const long long e1 = min(it.n, e0 + kOptChunk);
    const bool vec_arena = (it.reserved & 1) != 0;
    const bool sn = it.u != nullptr;
    const int W = pc.world;
    const bool use_mc = (vec_arena ? pc.vmc : pc.wmc) != nullptr;
    float* gmine = const_cast<float*>(it.g);
    float acc = 0.f, z = 0.f;
    const bool direct = !sn || (it.k == 1 && !it.flip && it.Cin_p == it.Cin);
    if (direct && (it.n & 3) == 0 && aligned16(it.g, it.p)) {
        for (long long e = e0 + threadIdx.x * 4; e < e1; e += kOptThreads * 4) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_mc) {
                g = multimem_sum4(mc_grad(pc, it.g + e, vec_arena));
            } else {
                for (int r = 0; r < W; ++r) {
                    const float4 v = *reinterpret_cast<const float4*>(peer_grad(pc, it.g + e, vec_arena, r));
                    g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
                }
            }
            *reinterpret_cast<float4*>(gmine + e) = g;
            if (sn) {
                const float4 w = *reinterpret_cast<const float4*>(it.p + e);
                acc += g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
            } else {
                z = fmaf(0.f, g.x + g.y + g.z + g.w, z);
            }
        }
    } else {
        for (long long e = e0 + threadIdx.x; e < e1; e += kOptThreads) {
            long long gi = e;
            if (sn) {
                int o, q;
                decode(it, (int)e, o, q, gi);
            }
            float g = 0.f;
            if (use_mc) {
                g = multimem_sum1(mc_grad(pc, it.g + gi, vec_arena));
            } else {
                for (int r = 0; r < W; ++r) g += *peer_grad(pc, it.g + gi, vec_arena, r);
            }
            gmine[gi] = g;
            if (sn) acc += g * it.p[e]; else z = fmaf(0.f, g, z);
        }
    }
    double t = block_sum((double)(sn ? acc : z), sh);
    if (threadIdx.x == 0) {
        if (sn) atomicAdd(it.dot, t);
        if (bad != nullptr && !isfinite(t)) atomicAdd(bad, 1.0);
    }
    }
}

struct AdamArgs {
    float lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale;
    int skip, pad;
};

// One thread between the two passes.  Without a scaler: bias corrections from the host's step counter.  With the
// device-resident dynamic loss scaler (fp16 operands: include/simulgen_b200.h, sg_scaler_state): a step whose gradients
// overflowed is SKIPPED (no parameter, moment or step-count change - torch.cuda.amp.GradScaler semantics) and the scale
// backs off; `growth_interval` clean steps in a row grow it.  Nothing here needs the host: the Trainer multiplies the
// loss by state->scale on the device and never reads it back.
__global__ void opt_prologue_kernel(AdamArgs a, int host_step, sg_scaler_state* __restrict__ st, const double* __restrict__ bad,
                                    AdamArgs* __restrict__ out, double* __restrict__ gnorm_sq) {
    int step = host_step;
    a.skip = 0;
    if (st != nullptr) {
        const float used = st->scale;
        a.grad_scale = a.grad_scale / used;
        if (bad != nullptr && bad[0] != 0.0) {
            a.skip = 1;
            st->scale = fmaxf(used * st->backoff, st->min_scale);
            st->good_steps = 0;
            st->skipped += 1;
            st->last_skipped = 1;
            if (gnorm_sq != nullptr) gnorm_sq[0] = (double)INFINITY;
        } else {
            st->step += 1;
            st->last_skipped = 0;
            if (++st->good_steps >= st->growth_interval) {
                st->scale = fminf(used * st->growth, st->max_scale);
                st->good_steps = 0;
            }
        }
        step = st->step;
    }
    a.bc1 = (float)(1.0 - pow((double)a.b1, (double)step));
    a.bc2_sqrt = sqrtf((float)(1.0 - pow((double)a.b2, (double)step)));
    *out = a;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamArgs& a) {
    p = p * (1.f - a.lr * a.wd);
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - (a.lr / a.bc1) * (m / denom);
}

__global__ void __launch_bounds__(kOptThreads)
opt_step_kernel(const sg_opt_item* __restrict__ items, const __grid_constant__ OptPrefix pf, const AdamArgs* __restrict__ args,
                double* __restrict__ gnorm_sq, const __grid_constant__ sg_peer pc) {
    __shared__ double sh[32];
    const AdamArgs a = *args;
    if (a.skip) return;
    // updated parameters go to this rank's buffer and, data parallel over peer memory, to every other rank's (fused
    // all-gather: world - 1 remote stores per element over NVLink)
    const bool mc_st = pc.world > 1 && pc.pmc != nullptr;
    auto put4 = [&](float* dst, const float4& val) {
        if (mc_st) { multimem_st4(mc_param(pc, dst), val); return; }        // lands in every rank's copy, mine included
        *reinterpret_cast<float4*>(dst) = val;
        for (int r = 0; r < pc.world; ++r)
            if (r != pc.rank) *reinterpret_cast<float4*>(peer_param(pc, dst, r)) = val;
    };
    auto put1 = [&](float* dst, float val) {
        if (mc_st) { multimem_st1(mc_param(pc, dst), val); return; }
        *dst = val;
        for (int r = 0; r < pc.world; ++r)
            if (r != pc.rank) *peer_param(pc, dst, r) = val;
    };
    for (int chunk = blockIdx.x; chunk < pf.start[pf.n_items]; chunk += gridDim.x) {      // grid-stride, see peer_reduce_dot_kernel
    const int idx = find_item(pf, chunk);
    const sg_opt_item it = items[idx];
    const long long e0 = (long long)(chunk - pf.start[idx]) * kOptChunk;
    const long long e1 = min(it.n, e0 + kOptChunk);
    const bool sn = it.u != nullptr;
    float inv_sigma = 1.f, coef = 0.f;
    if (sn) {
        float s = it.sigma[0];
        inv_sigma = 1.f / s;
        coef = (float)(it.dot[0] / (double)s);
    }
    float ss = 0.f;
    const bool direct = !sn || (it.k == 1 && !it.flip && it.Cin_p == it.Cin);
    if (direct && (it.n & 3) == 0 && (!sn || ((it.Cin & 3) == 0 && aligned16(it.vv, it.vv))) && aligned16(it.g, it.p) &&
        aligned16(it.m, it.v)) {
        for (long long e = e0 + threadIdx.x * 4; e < e1; e += kOptThreads * 4) {
            float4 g4 = *reinterpret_cast<const float4*>(it.g + e);
            float4 p4 = *reinterpret_cast<const float4*>(it.p + e);
            float4 m4 = *reinterpret_cast<const float4*>(it.m + e);
            float4 v4 = *reinterpret_cast<const float4*>(it.v + e);
            float g[4] = {g4.x, g4.y, g4.z, g4.w};
            float p[4] = {p4.x, p4.y, p4.z, p4.w};
            float m[4] = {m4.x, m4.y, m4.z, m4.w};
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
            if (sn) {
                int o = (int)(e / it.Cin);
                int q = (int)(e - (long long)o * it.Cin);
                float cu = coef * __ldg(it.u + o);
                float4 vv = __ldg(reinterpret_cast<const float4*>(it.vv + q));
                g[0] = (g[0] - cu * vv.x) * inv_sigma;
                g[1] = (g[1] - cu * vv.y) * inv_sigma;
                g[2] = (g[2] - cu * vv.z) * inv_sigma;
                g[3] = (g[3] - cu * vv.w) * inv_sigma;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float gi = g[i] * a.grad_scale;
                ss += gi * gi;
                adam_update(p[i], m[i], v[i], gi, a);
            }
            put4(it.p + e, make_float4(p[0], p[1], p[2], p[3]));
            *reinterpret_cast<float4*>(it.m + e) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(it.v + e) = make_float4(v[0], v[1], v[2], v[3]);
        }
    } else if (sn && !it.flip && (it.n & 3) == 0 && aligned16(it.p, it.m) && aligned16(it.v, it.v)) {
        for (long long e = e0 + threadIdx.x * 4; e < e1; e += kOptThreads * 4) {
            float4 p4 = *reinterpret_cast<const float4*>(it.p + e);
            float4 m4 = *reinterpret_cast<const float4*>(it.m + e);
            float4 v4 = *reinterpret_cast<const float4*>(it.v + e);
            float p[4] = {p4.x, p4.y, p4.z, p4.w};
            float m[4] = {m4.x, m4.y, m4.z, m4.w};
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
            float g[4];
            int o, q;
            long long gi;
            decode(it, (int)e, o, q, gi);
            int i = q / it.k, j = q - i * it.k;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float graw = __ldg(it.g + ((long long)j * it.Cout + o) * it.Cin_p + i);
                g[t] = (graw - coef * __ldg(it.u + o) * __ldg(it.vv + i * it.k + j)) * inv_sigma;
                if (++j == it.k) { j = 0; if (++i == it.Cin) { i = 0; ++o; } }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float gi2 = g[t] * a.grad_scale;
                ss += gi2 * gi2;
                adam_update(p[t], m[t], v[t], gi2, a);
            }
            put4(it.p + e, make_float4(p[0], p[1], p[2], p[3]));
            *reinterpret_cast<float4*>(it.m + e) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(it.v + e) = make_float4(v[0], v[1], v[2], v[3]);
        }
    } else {
        for (long long e = e0 + threadIdx.x; e < e1; e += kOptThreads) {
            float g;
            if (sn) {
                int o, q;
                long long gi;
                decode(it, (int)e, o, q, gi);
                g = (it.g[gi] - coef * __ldg(it.u + o) * __ldg(it.vv + q)) * inv_sigma;
            } else {
                g = it.g[e];
            }
            g *= a.grad_scale;
            ss += g * g;
            float p = it.p[e], m = it.m[e], v = it.v[e];
            adam_update(p, m, v, g, a);
            put1(it.p + e, p);
            it.m[e] = m;
            it.v[e] = v;
        }
    }
    if (gnorm_sq != nullptr) {
        double t = block_sum((double)ss, sh);
        if (threadIdx.x == 0) atomicAdd(gnorm_sq, t);
    }
    }
}

}  // namespace sg

using namespace sg;

static int build_prefix(const sg_opt_item* items_host, int n_items, OptPrefix& pf, long long& total, bool& any_sn, const char* who) {
    SG_REQUIRE(n_items > 0 && n_items <= kMaxItems, "%s: n_items=%d out of range (max %d)", who, n_items, kMaxItems);
    pf.n_items = n_items;
    total = 0;
    any_sn = false;
    for (int i = 0; i < n_items; ++i) {
        const sg_opt_item& it = items_host[i];
        SG_REQUIRE(it.n > 0 && it.n < (1LL << 31), "%s: item %d has n=%lld", who, i, it.n);
        if (it.u != nullptr) {
            any_sn = true;
            // n is the number of elements of THIS item: the whole tensor, or - sharded optimiser - whole rows of it
            SG_REQUIRE(it.dot != nullptr && it.sigma != nullptr && it.vv != nullptr && it.k >= 1 && it.Cin_p >= it.Cin &&
                           it.n <= (long long)it.Cout * it.Cin * it.k &&
                           it.n % ((long long)(it.flip ? it.Cout : it.Cin) * it.k) == 0,
                       "%s: item %d has inconsistent spectral-norm fields", who, i);
        }
        pf.start[i] = (int)total;
        total += cdiv(it.n, kOptChunk);
    }
    SG_REQUIRE(total < (1LL << 31), "%s: too many chunks", who);
    pf.start[n_items] = (int)total;
    return 0;
}

static int check_peer(const sg_peer* peer, const char* who) {
    SG_REQUIRE(peer->world >= 1 && peer->world <= SG_MAX_PEERS && peer->rank >= 0 && peer->rank < peer->world,
               "%s: bad peer context (world %d, rank %d)", who, peer->world, peer->rank);
    for (int r = 0; r < peer->world; ++r)
        SG_REQUIRE(peer->wbase[r] != nullptr && peer->vbase[r] != nullptr && peer->pbase[r] != nullptr,
                   "%s: missing base pointer of rank %d", who, r);
    return 0;
}

extern "C" int sg_peer_reduce_dot(const sg_opt_item* items_dev, const sg_opt_item* items_host, int n_items, double* dots,
                                  int n_dots, int want_bad_flag, int clear_dots, int max_blocks, const sg_peer* peer,
                                  void* stream) {
    SG_REQUIRE(peer != nullptr && dots != nullptr && n_dots >= 6, "peer_reduce_dot: missing arguments");
    if (check_peer(peer, "peer_reduce_dot")) return 1;
    cudaStream_t st = as_stream(stream);
    OptPrefix pf;
    long long total;
    bool any_sn;
    if (build_prefix(items_host, n_items, pf, total, any_sn, "peer_reduce_dot")) return 1;
    if (clear_dots) cudaMemsetAsync(dots, 0, sizeof(double) * n_dots, st);
    double* bad = dots + (n_dots - 6);
    const unsigned grid = (unsigned)((max_blocks > 0 && max_blocks < total) ? max_blocks : total);
    peer_reduce_dot_kernel<<<grid, kOptThreads, 0, st>>>(items_dev, pf, *peer, want_bad_flag ? bad : nullptr);
    return check_launch("peer_reduce_dot");
}

extern "C" int sg_opt_step(const sg_opt_item* items_dev, const sg_opt_item* items_host, int n_items, double* dots,
                           int n_dots, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                           float grad_scale, double* gnorm_sq, sg_scaler_state* scaler, const sg_peer* peer, int phase,
                           int max_blocks, void* stream) {
    // dots[0 .. n_dots - 6): one <G, W> per spectral-norm layer; then 1 double "non-finite gradient seen" and
    // sizeof(AdamArgs) = 40 bytes (5 doubles) for the device copy of the step's scalars
    SG_REQUIRE(dots != nullptr && n_dots >= 6, "opt_step: dots buffer needs >= 6 trailing scratch elements");
    SG_REQUIRE(phase == 0 || phase == 2 || phase == 3, "opt_step: phase must be 0, 2 or 3");
    double* bad = dots + (n_dots - 6);
    AdamArgs* dev_args = reinterpret_cast<AdamArgs*>(dots + (n_dots - 5));
    cudaStream_t st = as_stream(stream);
    OptPrefix pf;
    long long total;
    bool any_sn;
    if (build_prefix(items_host, n_items, pf, total, any_sn, "opt_step")) return 1;
    sg_peer pc{};
    pc.world = 1;
    if (peer != nullptr) {
        if (check_peer(peer, "opt_step")) return 1;
        pc = *peer;
    }
    if (phase == 0) {
        cudaMemsetAsync(dots, 0, sizeof(double) * n_dots, st);
        if (any_sn || scaler != nullptr)
            opt_dot_kernel<<<(unsigned)total, kOptThreads, 0, st>>>(items_dev, pf, scaler != nullptr ? bad : nullptr);
    }
    if (phase != 3) {
        AdamArgs a{};
        a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.grad_scale = grad_scale;
        opt_prologue_kernel<<<1, 1, 0, st>>>(a, step, scaler, scaler != nullptr ? bad : nullptr, dev_args, gnorm_sq);
    }
    const unsigned grid = (unsigned)((max_blocks > 0 && max_blocks < total) ? max_blocks : total);
    opt_step_kernel<<<grid, kOptThreads, 0, st>>>(items_dev, pf, dev_args, gnorm_sq, pc);
    return check_launch("opt_step");
}
