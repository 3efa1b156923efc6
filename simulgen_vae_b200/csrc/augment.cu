// Batch assembly + on-the-fly augmentation in one pass.
//
// Replaces, for a GPU-resident dataset, AugmentedDataset.__getitem__ + default_collate of the reference
// (modules/augmentation.py:43-124, modules/utils.py:56-66): per sample of the batch
//     s = data[idx]                                   gather
//     s = s + noise_level * eps                        augmentation.py:86-89   (noise_level = 0: skipped)
//     s = s * scale                                    augmentation.py:91-95   (scale = 1: identity)
//     s = lam * s + (1 - lam) * data[other]            augmentation.py:114-124 (other < 0: skipped)
// with the products and sums rounded separately (no FMA contraction), i.e. bit-identical to the ATen sequence;
// (1 - lam) is formed by the host in double precision like the reference's Python scalar arithmetic.
// in that order (augmentation.py:57-84; shift and cutout have probability 0 in the reference).  The reference does
// each step as a separate ATen pass over a [N, T] sample plus a stacking copy; here every element is read once
// (twice with mixup) and written once: 8 B per element without mixup, 12 B with.  Optionally the bf16 operand of
// the first encoder conv ([N][B][Tp], zero tail) is written in the same pass, which removes sg_pack_input.
// eps comes from the counter-based Philox generator keyed on (seed, draw, dataset index, element) or, for the
// parity tests, from an injected tensor.
#include "common.cuh"

namespace sg {

constexpr int kAugWarps = 8;

template <bool VEC>
__global__ void __launch_bounds__(kAugWarps * 32)
assemble_batch_kernel(const float* __restrict__ data, const int* __restrict__ ids, const float* __restrict__ table,
                      const float* __restrict__ inj, float* __restrict__ out, __nv_bfloat16* __restrict__ op, int B, int N,
                      int T, int Tp, uint64_t seed, uint64_t draw) {
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * N, wstride = (long long)gridDim.x * kAugWarps;
    const size_t sample_elems = (size_t)N * T;
    for (long long row = (long long)blockIdx.x * kAugWarps + (threadIdx.x >> 5); row < rows; row += wstride) {
        const int n = (int)(row / B), b = (int)(row - (long long)n * B);      // b fastest: the operand rows (n, b) are contiguous
        const int idx = ids[b], oth = ids[B + b];
        const float nl = table[b], sc = table[B + b], lam = table[2 * B + b], om = table[3 * B + b];
        const float* src = data + (size_t)idx * sample_elems + (size_t)n * T;
        const float* src2 = oth >= 0 ? data + (size_t)oth * sample_elems + (size_t)n * T : nullptr;
        const float* nz = (inj != nullptr) ? inj + (size_t)b * sample_elems + (size_t)n * T : nullptr;
        float* dst = out != nullptr ? out + (size_t)b * sample_elems + (size_t)n * T : nullptr;
        for (int seg = lane; seg * 8 < Tp; seg += 32) {
            const int t0 = seg * 8;
            F8 v;
            if (t0 < T) {
                if (VEC && t0 + 8 <= T) {
                    float4 a = __ldcs(reinterpret_cast<const float4*>(src + t0)), c = __ldcs(reinterpret_cast<const float4*>(src + t0 + 4));
                    v.v[0] = a.x; v.v[1] = a.y; v.v[2] = a.z; v.v[3] = a.w; v.v[4] = c.x; v.v[5] = c.y; v.v[6] = c.z; v.v[7] = c.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = t0 + i < T ? __ldg(src + t0 + i) : 0.f;
                }
                if (nl != 0.f) {
                    float e[8];
                    if (nz != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] = t0 + i < T ? __ldg(nz + t0 + i) : 0.f;
                    } else {
                        float q0[4], q1[4];
                        const uint64_t quad = (uint64_t)n * (uint64_t)(Tp >> 2) + (uint64_t)(t0 >> 2);   // unique per (n, segment)
                        philox_normal4(seed, draw, (uint64_t)idx, quad, q0);
                        philox_normal4(seed, draw, (uint64_t)idx, quad + 1, q1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { e[i] = q0[i]; e[4 + i] = q1[i]; }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = __fadd_rn(v.v[i], __fmul_rn(e[i], nl));   // sample + randn * level, unfused like ATen
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] *= sc;
                if (src2 != nullptr) {
                    F8 w;
                    if (VEC && t0 + 8 <= T) {
                        float4 a = __ldcs(reinterpret_cast<const float4*>(src2 + t0)), c = __ldcs(reinterpret_cast<const float4*>(src2 + t0 + 4));
                        w.v[0] = a.x; w.v[1] = a.y; w.v[2] = a.z; w.v[3] = a.w; w.v[4] = c.x; w.v[5] = c.y; w.v[6] = c.z; w.v[7] = c.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w.v[i] = t0 + i < T ? __ldg(src2 + t0 + i) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = __fadd_rn(__fmul_rn(lam, v.v[i]), __fmul_rn(om, w.v[i]));
                }
                if (dst == nullptr) {
                    // operand-only batch (engine.PackedBatch): the fp32 copy is never materialised
                } else if (VEC && t0 + 8 <= T) {
                    *reinterpret_cast<float4*>(dst + t0) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
                    *reinterpret_cast<float4*>(dst + t0 + 4) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (t0 + i < T) dst[t0 + i] = v.v[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (t0 + i >= T) v.v[i] = 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] = 0.f;
            }
            if (op != nullptr) store8(op + (size_t)row * Tp + t0, v);
        }
    }
}

// Short rows (Tp == 8: static fields, T = 1 .. 8).  A warp per row would keep one lane of 32 busy over B * N rows.
// Tiles of 32 channels x 32 samples instead: phase 1 walks the 32 * T contiguous floats each sample has in the tile
// (lanes along (n, t): gather, augmentation and the fp32 store are coalesced) and leaves the values in shared memory,
// phase 2 writes the operand rows (n, b) from there, lanes along b (16 contiguous bytes per lane).  The Philox keying
// is the general kernel's: (seed, draw, dataset index, quad = n * Tp / 4 + t / 4), component t % 4.
constexpr int kAugShortTile = 32;
constexpr int kAugShortPitch = kAugShortTile * 8 + 1;

__global__ void __launch_bounds__(kAugWarps * 32)
assemble_batch_short_kernel(const float* __restrict__ data, const int* __restrict__ ids, const float* __restrict__ table,
                            const float* __restrict__ inj, float* __restrict__ out, __nv_bfloat16* __restrict__ op, int B,
                            int N, int T, uint64_t seed, uint64_t draw) {
    __shared__ float xs[kAugShortTile][kAugShortPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t sample_elems = (size_t)N * T;
    // without an operand to write there is no transposition: the tile widens to 1024 channels (4 KB runs per sample at T = 1
    // instead of 128-byte pieces) and shared memory is not used
    const int tile_n = op != nullptr ? kAugShortTile : 1024;
    const int tiles_b = (B + kAugShortTile - 1) / kAugShortTile;
    const long long tiles = (long long)((N + tile_n - 1) / tile_n) * tiles_b;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n0 = (int)(tile / tiles_b) * tile_n, b0 = (int)(tile % tiles_b) * kAugShortTile;
        const int run = min(tile_n, N - n0) * T;
        __syncthreads();
        for (int bl = warp; bl < kAugShortTile && b0 + bl < B; bl += kAugWarps) {
            const int b = b0 + bl;
            const int idx = ids[b], oth = ids[B + b];
            const float nl = table[b], sc = table[B + b], lam = table[2 * B + b], om = table[3 * B + b];
            const size_t base = (size_t)n0 * T;
            const float* src = data + (size_t)idx * sample_elems + base;
            const float* src2 = oth >= 0 ? data + (size_t)oth * sample_elems + base : nullptr;
            const float* nz = inj != nullptr ? inj + (size_t)b * sample_elems + base : nullptr;
            float* dst = out != nullptr ? out + (size_t)b * sample_elems + base : nullptr;
            for (int i = lane; i < run; i += 32) {
                float v = __ldcs(src + i);
                if (nl != 0.f) {
                    float e;
                    if (nz != nullptr) {
                        e = __ldg(nz + i);
                    } else {
                        const int n = n0 + i / T, t = i % T;
                        float q[4];
                        philox_normal4(seed, draw, (uint64_t)idx, (uint64_t)n * 2 + (uint64_t)(t >> 2), q);
                        e = (t & 2) ? ((t & 1) ? q[3] : q[2]) : ((t & 1) ? q[1] : q[0]);
                    }
                    v = __fadd_rn(v, __fmul_rn(e, nl));
                }
                v = __fmul_rn(v, sc);
                if (src2 != nullptr) v = __fadd_rn(__fmul_rn(lam, v), __fmul_rn(om, __ldcs(src2 + i)));
                if (dst != nullptr) dst[i] = v;
                if (op != nullptr) xs[bl][i] = v;
            }
        }
        if (op == nullptr) continue;
        __syncthreads();
        const int b = b0 + lane;
#pragma unroll
        for (int j = 0; j < kAugShortTile / kAugWarps; ++j) {
            const int nloc = warp + j * kAugWarps, n = n0 + nloc;
            if (n < N && b < B) {
                F8 r;
#pragma unroll
                for (int i = 0; i < 8; ++i) r.v[i] = i < T ? xs[lane][nloc * T + i] : 0.f;
                store8(op + ((size_t)n * B + b) * 8, r);
            }
        }
    }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_assemble_batch(const float* data, int P, const int* ids, const float* table, const float* injected_noise,
                                 float* out, void* operand, int B, int N, int T, int Tp, unsigned long long seed,
                                 unsigned long long draw, int blocks_per_sm, void* stream) {
    SG_REQUIRE(B > 0 && N > 0 && T > 0 && P > 0, "assemble_batch: bad shape");
    SG_REQUIRE(out != nullptr || operand != nullptr, "assemble_batch: neither an fp32 batch nor an operand to write");
    SG_REQUIRE(operand == nullptr || (Tp % 8 == 0 && Tp >= T), "assemble_batch: bad Tp=%d for T=%d", Tp, T);
    if (operand == nullptr) Tp = (T + 7) / 8 * 8;
    const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    // blocks_per_sm (0 = 16): a prefetching loader runs this kernel on its own stream UNDERNEATH the training step; a
    // small resident footprint (2 blocks = 512 threads per SM) leaves the thread slots and registers the persistent
    // tensor-core GEMM CTAs need to start next to it - with the full grid the SMs fill up with these blocks first and
    // the step waits for the whole gather (measured: resident e2e 0.906 of the device-resident rate, r2_bench_full_d.json)
    if (blocks_per_sm <= 0 || blocks_per_sm > 16) blocks_per_sm = 16;
    long long blocks = cdiv((long long)B * N, kAugWarps);
    int grid = (int)(blocks < 148LL * blocks_per_sm ? blocks : 148LL * blocks_per_sm);
    cudaStream_t st = as_stream(stream);
    if (Tp == 8 && (reinterpret_cast<uintptr_t>(operand) & 15) == 0) {
        const long long tiles = cdiv((long long)N, operand != nullptr ? kAugShortTile : 1024) * cdiv((long long)B, kAugShortTile);
        int nb = 0;                                           // the grid is persistent: launch exactly what is resident
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, assemble_batch_short_kernel, kAugWarps * 32, 0) != cudaSuccess || nb < 1) nb = 1;
        const long long cap = 148LL * (blocks_per_sm < nb ? blocks_per_sm : nb);
        assemble_batch_short_kernel<<<(int)(tiles < cap ? tiles : cap), kAugWarps * 32, 0, st>>>(
            data, ids, table, injected_noise, out, (__nv_bfloat16*)operand, B, N, T, seed, draw);
        return check_launch("assemble_batch");
    }
    if (vec)
        assemble_batch_kernel<true><<<grid, kAugWarps * 32, 0, st>>>(data, ids, table, injected_noise, out, (__nv_bfloat16*)operand,
                                                                     B, N, T, Tp, seed, draw);
    else
        assemble_batch_kernel<false><<<grid, kAugWarps * 32, 0, st>>>(data, ids, table, injected_noise, out, (__nv_bfloat16*)operand,
                                                                      B, N, T, Tp, seed, draw);
    return check_launch("assemble_batch");
}
