// CTA-pair (cta_group::2) variant of the tcgen05 implicit-GEMM kernel.  Included by gemm_tc.cu.
//
// Two CTAs of a cluster (same TPC) work on one 256 x 256 output tile: CTA r owns rows [128 r, 128 r + 128)
// of A and of the accumulator, and loads only HALF of the B tile (columns [128 r, 128 r + 128)); the
// tcgen05.mma.cta_group::2 issued by the leader reads A and B from both CTAs' shared memory.  Per CTA and
// 64-deep k-block that is 32 KB of TMA traffic instead of 48 KB for the same FLOPs: the single-CTA kernel
// was bound by L2 -> SM bandwidth / power (profiles/r1_ncu_full_gemm_*: L2 throughput 55-68 %, tensor pipe
// 45-75 % active, SM clock capped at 1.2-1.5 GHz), so bytes per FLOP is the lever.
//
// Protocol (per smem stage s, per accumulator buffer a):
//   full[s]   (leader CTA, count 1): leader's producer arrives with expect_tx(64 KB); the TMA loads of BOTH
//             CTAs complete_tx on it (the peer's loads name the leader's barrier: address bit 24 cleared)
//   empty[s]  (each CTA, count 1):   tcgen05.commit ... multicast::cluster from the leader's MMA thread
//   tfull[a]  (each CTA, count 1):   same multicast commit after the last k-block of a tile
//   tempty[a] (leader CTA, count 2 x EPI_WARPS): the epilogue warps of both CTAs arrive (the peer's remotely)
//
// Epilogue: TMEM -> registers (tcgen05.ld 32x32b.x32: lane = row) -> per-warp padded smem slab -> global
// stores in which 8 consecutive lanes write one 128-byte row segment (the single-CTA kernel stores 16 bytes
// per lane into 32 different rows per instruction).
#pragma once

namespace sg {
namespace pair {

constexpr int BMH = 128;            // rows per CTA
constexpr int PM = 256;             // rows per pair tile
constexpr int BN = 256, BNH = 128, BK = 64, STAGES = 6;
constexpr int BOX_BYTES = 64 * 64 * 2;
constexpr int A_BYTES = BMH * BK * 2;                 // 16 KB
constexpr int B_BYTES = BNH * BK * 2;                 // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 32 KB
constexpr int EPI_PITCH = 33;                         // floats per staged row (bank-conflict-free)
#ifndef SG_PAIR_EPI_WARPS
#define SG_PAIR_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = SG_PAIR_EPI_WARPS;          // 4: one warp per TMEM lane quarter; 8: two, splitting the columns
constexpr int EPI_CHUNKS = (BN / 32) * 4 / EPI_WARPS; // 32-column chunks per warp and tile
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
static_assert(SMEM_BYTES <= 232448, "pair kernel shared memory");
constexpr int TMEM_COLS = 512;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint32_t bar_addr, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same smem offset in the leader CTA (rank 0 of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

struct PairWork {
    int m0, n0, z, it_lo, it_hi, split;
};

// tile order: groups of `group_m` m-tiles; inside a group n is the slow index, so one wave of CTA pairs covers a
// roughly square patch of the output and every operand row/column it needs is fetched from DRAM once per wave.
__device__ __forceinline__ PairWork decode_pair_work(const TcParams& p, int w, int total_iters) {
    PairWork r;
    int tiles = p.m_tiles * p.n_tiles;
    int t = w % tiles;
    int rest = w / tiles;
    r.z = rest % p.z_count;
    r.split = rest / p.z_count;
    int per_group = p.group_m * p.n_tiles;
    int g = t / per_group;
    int in_g = t - g * per_group;
    int gm = min(p.group_m, p.m_tiles - g * p.group_m);     // m-tiles in this (possibly last, smaller) group
    int nt = in_g / gm, mt = g * p.group_m + (in_g - nt * gm);
    r.m0 = mt * PM;
    r.n0 = nt * BN;
    int per = (total_iters + p.splits - 1) / p.splits;
    r.it_lo = r.split * per;
    r.it_hi = min(total_iters, r.it_lo + per);
    return r;
}

// One tile of the epilogue for one warp: TMEM lane quarter q (32 rows, lane = row) x BN columns, in chunks of 32 columns.
// The tcgen05.ld of chunk c + 1 is in flight while chunk c is processed.  fp32 output: the chunk is staged through a
// padded smem slab and leaves as 128-byte row segments (plain, accumulate or red.add).  16-bit output (OUT16): values are
// packed before staging and two chunks (64 columns = 128 bytes per row) leave together as 16-byte stores.
template <int MODE, bool OUT16>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, const PairWork& wk, uint32_t tmem_acc, float* slab,
                                              int m_base, int lane, int c_lo) {
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    float bias = 0.f;
    if (p.bias != nullptr && m_base + lane < p.M && wk.split == 0) bias = p.bias[m_base + lane];
    const bool have_k = wk.it_hi > wk.it_lo;
    float* obase = p.out + (long long)wk.z * p.c_sz;
    uint32_t* slab_u = reinterpret_cast<uint32_t*>(slab);
    // GroupNorm statistics of the tile (fprop, no split-K): this lane's row, running sums of the sample the current
    // columns belong to; flushed with two atomics whenever the sample changes and at the tile end
    const bool do_stats = (MODE == MODE_FPROP) && p.rowstat != nullptr;
    float st_s = 0.f, st_ss = 0.f;
    int st_b = do_stats ? (wk.n0 + c_lo * 32) / p.st_Tp : 0;
    auto st_flush = [&](int b) {
        if (m_base + lane < p.M && b < p.st_B) {
            float* rs = p.rowstat + ((size_t)b * p.M + (m_base + lane)) * 2;
            atomicAdd(rs, st_s);
            atomicAdd(rs + 1, st_ss);
        }
        st_s = 0.f;
        st_ss = 0.f;
    };
    auto chunk = [&](uint32_t* v, int c) {
        const int col0 = wk.n0 + c * 32;
        if (have_k && col0 < p.N) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + bias);
            if (do_stats) {
                int b = col0 / p.st_Tp, t = col0 - b * p.st_Tp;
                if (b != st_b) { st_flush(st_b); st_b = b; }
                if (t + 32 <= p.st_T && col0 + 32 <= p.N) {
                    // the usual case: all 32 columns are valid time steps of one sample
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float val = __uint_as_float(v[e]);
                        st_s += val;
                        st_ss = fmaf(val, val, st_ss);
                    }
                } else if (p.st_Tp >= 32) {
                    // at most one sample boundary inside the 32 columns: [0, nb) -> sample b, [nb, 32) -> b + 1
                    const int nb = p.st_Tp - t;
                    const int lim0 = min(min(nb, p.st_T - t), p.N - col0);
                    const int lim1 = min(nb + p.st_T, p.N - col0);
                    float s1 = 0.f, ss1 = 0.f;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float val = __uint_as_float(v[e]);
                        if (e < lim0) { st_s += val; st_ss = fmaf(val, val, st_ss); }
                        if (e >= nb && e < lim1) { s1 += val; ss1 = fmaf(val, val, ss1); }
                    }
                    if (nb < 32) {
                        st_flush(st_b);
                        st_b = b + 1;
                        st_s = s1;
                        st_ss = ss1;
                    }
                } else {
                    // short rows (static fields, Tp = 8): several samples per chunk
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float val = __uint_as_float(v[e]);
                        if (t < p.st_T && col0 + e < p.N) { st_s += val; st_ss = fmaf(val, val, st_ss); }
                        if (++t == p.st_Tp) { st_flush(st_b); t = 0; st_b = st_b + 1; }
                    }
                }
            }
            if (OUT16) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    slab_u[lane * EPI_PITCH + (c & 1) * 16 + e] =
                        f2_to_op16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) slab_u[lane * EPI_PITCH + e] = v[e];
                __syncwarp();
                const int n = col0 + sub_c;
                if (n < p.N) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rl = i * 4 + sub_r;
                        const int m = m_base + rl;
                        if (m < p.M) {
                            const float* sp = slab + rl * EPI_PITCH + sub_c;
                            float4 r = make_float4(sp[0], sp[1], sp[2], sp[3]);
                            float* dst = obase + (long long)m * p.ldc + n;
                            if (p.atomic) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(r.x), "f"(r.y),
                                             "f"(r.z), "f"(r.w)
                                             : "memory");
                            } else {
                                if (p.accumulate) {
                                    float4 o = *reinterpret_cast<const float4*>(dst);
                                    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
                                }
                                *reinterpret_cast<float4*>(dst) = r;
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (OUT16 && (c & 1)) {
            // two staged chunks: 32 rows x 64 columns of 16-bit values, 8 lanes x 16 bytes per row
            const int cbase = col0 - 32;
            if (have_k && cbase < p.N) {
                __syncwarp();
                const int n = cbase + sub_c * 2;
                if (n < p.N) {
                    __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(obase);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rl = i * 4 + sub_r;
                        const int m = m_base + rl;
                        if (m < p.M) {
                            const uint32_t* sp = slab_u + rl * EPI_PITCH + sub_c;
                            uint4 nv = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                            uint4* dst = reinterpret_cast<uint4*>(o16 + (long long)m * p.ldc + n);
                            if (p.accumulate) {
                                // 16-bit destination that already holds a gradient contribution: add in fp32, round once more
                                const uint4 ov = *dst;
                                const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ov);
                                uint32_t* nw = reinterpret_cast<uint32_t*>(&nv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 a = op16x2_to_f2(nw[e]), b = op16x2_to_f2(ow[e]);
                                    nw[e] = f2_to_op16x2(a.x + b.x, a.y + b.y);
                                }
                            }
                            *dst = nv;
                        }
                    }
                }
                __syncwarp();
            }
        }
    };
    uint32_t va[32], vb[32];
    tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(c_lo * 32), va);
#pragma unroll 1
    for (int c = c_lo; c < c_lo + EPI_CHUNKS; c += 2) {
        tmem_ld_wait_regs(va);
        tmem_ld_32x32b_x32(tmem_acc + (uint32_t)((c + 1) * 32), vb);
        chunk(va, c);
        tmem_ld_wait_regs(vb);
        if (c + 2 < c_lo + EPI_CHUNKS) tmem_ld_32x32b_x32(tmem_acc + (uint32_t)((c + 2) * 32), va);
        chunk(vb, c + 1);
    }
    if (do_stats) st_flush(st_b);
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    constexpr bool A_MN = (MODE == MODE_DGRAD);
    constexpr bool B_MN = (MODE != MODE_WGRAD);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_iters = p.taps * p.kblocks;
    const int num_work = p.m_tiles * p.n_tiles * p.z_count * p.splits;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = pair_id; w < num_work; w += num_pairs) {
                PairWork wk = decode_pair_work(p, w, total_iters);
                const int m_cta = wk.m0 + (int)rank * BMH;
                const int n_cta = wk.n0 + (int)rank * BNH;
                for (int it = wk.it_lo; it < wk.it_hi; ++it) {
                    int kb = it / p.taps, j = it - kb * p.taps;   // taps innermost: shifted reloads hit L2
                    int k0 = kb * BK;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    const uint32_t bar = smem_u32(&full_bar[stage]) & kPeerMask;     // the leader's barrier
                    if (MODE == MODE_FPROP) {
                        tma_load_3d_2sm(&tmA, bar, sa, k0, m_cta, j);
                        tma_load_3d_2sm(&tmA, bar, sa + BOX_BYTES, k0, m_cta + 64, j);
                        const int pl = p.b_plane0 + j * p.b_plane_step;
                        tma_load_3d_2sm(&tmB, bar, sb, n_cta, k0, pl);
                        tma_load_3d_2sm(&tmB, bar, sb + BOX_BYTES, n_cta + 64, k0, pl);
                    } else if (MODE == MODE_DGRAD) {
                        tma_load_3d_2sm(&tmA, bar, sa, m_cta, k0, j);
                        tma_load_3d_2sm(&tmA, bar, sa + BOX_BYTES, m_cta + 64, k0, j);
                        const int pl = p.b_plane0 + j * p.b_plane_step;
                        tma_load_3d_2sm(&tmB, bar, sb, n_cta, k0, pl);
                        tma_load_3d_2sm(&tmB, bar, sb + BOX_BYTES, n_cta + 64, k0, pl);
                    } else {
                        tma_load_3d_2sm(&tmA, bar, sa, k0, m_cta, p.a_plane);
                        tma_load_3d_2sm(&tmA, bar, sa + BOX_BYTES, k0, m_cta + 64, p.a_plane);
                        const int pl = p.b_plane0 + wk.z * p.b_plane_step;
                        tma_load_3d_2sm(&tmB, bar, sb, k0, n_cta, pl);
                        tma_load_3d_2sm(&tmB, bar, sb + BOX_BYTES, k0, n_cta + 64, pl);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();                   // reconverge: the cluster barrier below is .aligned
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, M=256 (pair), N=256, majors per mode
            const uint32_t idesc = (1u << 4) | ((uint32_t)p.a_fmt << 7) | ((uint32_t)p.b_fmt << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int tile_iter = 0;
            for (int w = pair_id; w < num_work; w += num_pairs, ++tile_iter) {
                PairWork wk = decode_pair_work(p, w, total_iters);
                int as = tile_iter & 1;
                uint32_t aphase = (tile_iter >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tcgen05_fence_after();
                uint32_t tmem_d = tmem_base + as * BN;
                for (int it = wk.it_lo; it < wk.it_hi; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    uint32_t sb = sa + A_BYTES;
                    uint64_t da = A_MN ? make_desc(sa, BOX_BYTES, 1024) : make_desc(sa, 16, 1024);
                    uint64_t db = B_MN ? make_desc(sb, BOX_BYTES, 1024) : make_desc(sb, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        uint64_t a = da + (uint64_t)((A_MN ? 2048 : 32) * kk >> 4);
                        uint64_t b = db + (uint64_t)((B_MN ? 2048 : 32) * kk >> 4);
                        tcgen05_mma_bf16_2sm(tmem_d, a, b, idesc, (it > wk.it_lo || kk > 0) ? 1u : 0u);
                    }
                    tcgen05_commit_2sm(&empty_bar[stage]);      // frees the slot in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit_2sm(&tfull_bar[as]);             // accumulator ready in both CTAs
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2.., both CTAs) =====================
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int c_lo = ((warp - 2) >> 2) * EPI_CHUNKS;     // first of this warp's 32-column chunks
        float* slab = epi + (warp - 2) * 32 * EPI_PITCH;
        int tile_iter = 0;
        for (int w = pair_id; w < num_work; w += num_pairs, ++tile_iter) {
            PairWork wk = decode_pair_work(p, w, total_iters);
            int as = tile_iter & 1;
            uint32_t aphase = (tile_iter >> 1) & 1;
            mbar_wait(&tfull_bar[as], aphase);
            tcgen05_fence_after();
            const int m_base = wk.m0 + (int)rank * BMH + q * 32;
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            if (p.out_bf16) epilogue_tile<MODE, true>(p, wk, tmem_acc, slab, m_base, lane, c_lo);
            else epilogue_tile<MODE, false>(p, wk, tmem_acc, slab, m_base, lane, c_lo);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty_bar[as]); else mbar_arrive_leader(&tempty_bar[as]);
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();                 // nobody leaves (or frees TMEM) while the peer may still touch this CTA
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

}  // namespace pair
}  // namespace sg
