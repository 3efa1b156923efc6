// tcgen05 / TMEM / TMA implicit-GEMM kernels for the stride-1 "same" Conv1d family (sm_100a).
//
// Replaces cuDNN fprop / dgrad / wgrad behind nn.Conv1d / nn.ConvTranspose1d in
// encoder.py:34,43, common.py:84,110,135-141, decoder.py:31,118,135,145,155,164 and their autograd
// backward (train.py:153).  Activations are in the CR layout [C][B*Tp] (r contiguous, zero gap
// between samples = the conv padding), weights in the GEMM layout Wg[k][Cout][Cin_p].
//
//   fprop  D[co][r]      = sum_{j,ci} Wg[j][co][ci] * act[ci][r + j - pad]   A: K-major,  B: MN-major
//   dgrad  D[ci][r]      = sum_{j,co} Wg[j][co][ci] * dy [co][r - j + pad]   A: MN-major, B: MN-major
//   wgrad  D[j][co][ci]  = sum_r      dy[co][r]     * act[ci][r + j - pad]   A: K-major,  B: K-major
//
// One persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM
// alloc), warps 2-5 = epilogue (TMEM -> registers -> global).  Tile 128 x 256 x 64, 4 smem stages
// (48 KB each, 128B-swizzled TMA boxes of 64 x 64 bf16), two 256-column fp32 accumulators in TMEM
// so the epilogue of tile i overlaps the main loop of tile i+1.  Out-of-bounds box elements are
// zero-filled by TMA (ragged edges need no special code).  TMA moves 16-byte granules, so a conv tap
// cannot be a 1-element shift of a box along the contiguous r axis: operands of k-tap convs are stored
// as k pre-shifted planes (common.cuh, store_row_planes) and tap j selects a plane through the third
// TMA coordinate - the conv's zero padding is already baked into the planes.
#include <stdlib.h>

#include "tc_common.cuh"

namespace sg {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int BOX_BYTES = 64 * 64 * 2;              // one 64 x 64 bf16 TMA box
constexpr int A_BYTES = BM * BK * 2;                // 16 KB  (2 boxes)
constexpr int B_BYTES = BN * BK * 2;                // 32 KB  (4 boxes)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;

struct Work {
    int m0, n0, z, it_lo, it_hi, split;
};

__device__ __forceinline__ Work decode_work(const TcParams& p, int w, int total_iters) {
    Work r;
    int tiles = p.m_tiles * p.n_tiles;
    int t = w % tiles;
    int rest = w / tiles;
    r.z = rest % p.z_count;
    r.split = rest / p.z_count;
    int mt, nt;
    if (p.m_fastest) { mt = t % p.m_tiles; nt = t / p.m_tiles; } else { nt = t % p.n_tiles; mt = t / p.n_tiles; }
    r.m0 = mt * BM;
    r.n0 = nt * BN;
    int per = (total_iters + p.splits - 1) / p.splits;
    r.it_lo = r.split * per;
    r.it_hi = min(total_iters, r.it_lo + per);
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
            mbar_init(&tempty_bar[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_iters = p.taps * p.kblocks;
    const int num_work = p.m_tiles * p.n_tiles * p.z_count * p.splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                Work wk = decode_work(p, w, total_iters);
                for (int it = wk.it_lo; it < wk.it_hi; ++it) {
                    int kb = it / p.taps, j = it - kb * p.taps;   // taps innermost: shifted reloads hit L2
                    int k0 = kb * BK;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    if (MODE == MODE_FPROP) {
                        tma_load_3d(&tmA, &full_bar[stage], sa, k0, wk.m0, j);
                        tma_load_3d(&tmA, &full_bar[stage], sa + BOX_BYTES, k0, wk.m0 + 64, j);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_3d(&tmB, &full_bar[stage], sb + i * BOX_BYTES, wk.n0 + 64 * i, k0,
                                        p.b_plane0 + j * p.b_plane_step);
                    } else if (MODE == MODE_DGRAD) {
                        tma_load_3d(&tmA, &full_bar[stage], sa, wk.m0, k0, j);
                        tma_load_3d(&tmA, &full_bar[stage], sa + BOX_BYTES, wk.m0 + 64, k0, j);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_3d(&tmB, &full_bar[stage], sb + i * BOX_BYTES, wk.n0 + 64 * i, k0,
                                        p.b_plane0 + j * p.b_plane_step);
                    } else {
                        tma_load_3d(&tmA, &full_bar[stage], sa, k0, wk.m0, p.a_plane);
                        tma_load_3d(&tmA, &full_bar[stage], sa + BOX_BYTES, k0, wk.m0 + 64, p.a_plane);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            tma_load_3d(&tmB, &full_bar[stage], sb + i * BOX_BYTES, k0, wk.n0 + 64 * i,
                                        p.b_plane0 + wk.z * p.b_plane_step);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, M=128, N=256, majors per mode
            const uint32_t idesc = (1u << 4) | ((uint32_t)p.a_fmt << 7) | ((uint32_t)p.b_fmt << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int tile_iter = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++tile_iter) {
                Work wk = decode_work(p, w, total_iters);
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
                    // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); +32 B per UMMA_K=16
                    // MN-major: 64-element atoms 8192 B apart (LBO = one box), 8-k groups 1024 B apart (SBO);
                    //           +2048 B per UMMA_K=16
                    uint64_t da = A_MN ? make_desc(sa, BOX_BYTES, 1024) : make_desc(sa, 16, 1024);
                    uint64_t db = B_MN ? make_desc(sb, BOX_BYTES, 1024) : make_desc(sb, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        uint64_t a = da + (uint64_t)((A_MN ? 2048 : 32) * kk >> 4);
                        uint64_t b = db + (uint64_t)((B_MN ? 2048 : 32) * kk >> 4);
                        tcgen05_mma_bf16(tmem_d, a, b, idesc, (it > wk.it_lo || kk > 0) ? 1u : 0u);
                    }
                    tcgen05_commit(&empty_bar[stage]);      // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(&tfull_bar[as]);             // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        int tile_iter = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++tile_iter) {
            Work wk = decode_work(p, w, total_iters);
            int as = tile_iter & 1;
            uint32_t aphase = (tile_iter >> 1) & 1;
            mbar_wait(&tfull_bar[as], aphase);
            tcgen05_fence_after();
            const int m = wk.m0 + q * 32 + lane;
            const bool m_ok = m < p.M;
            float bias = 0.f;
            if (p.bias != nullptr && m_ok && wk.split == 0) bias = p.bias[m];
            float* orow = p.out + (long long)wk.z * p.c_sz + (long long)(m_ok ? m : 0) * p.ldc;
            const bool have_k = wk.it_hi > wk.it_lo;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32);
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                int n = wk.n0 + c * 32;
                if (m_ok && have_k) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        if (n + e < p.N) {
                            float4 r = make_float4(__uint_as_float(v[e]) + bias, __uint_as_float(v[e + 1]) + bias,
                                                   __uint_as_float(v[e + 2]) + bias, __uint_as_float(v[e + 3]) + bias);
                            float* dst = orow + n + e;
                            if (p.atomic) {
                                atomicAdd(dst, r.x); atomicAdd(dst + 1, r.y); atomicAdd(dst + 2, r.z); atomicAdd(dst + 3, r.w);
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
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// 3-D 16-bit operand [planes][rows][cols] (cols contiguous), box 64 x 64 x 1, 128B swizzle, zero OOB fill.
// The tensor-map data type is BFLOAT16 in BOTH library builds (bf16 and, with -DSG_OP16_HALF, fp16 operands): TMA only
// moves 2-byte elements here (no arithmetic, zero fill is all-zero bits in either format); the MMA's operand format is
// set by the instruction descriptor (TcParams::a_fmt / b_fmt).
static int make_map_op(CUtensorMap* m, const void* base, long long cols, long long rows, int planes,
                       long long plane_stride_elems) {
    EncodeTiledFn fn = get_encode_fn();
    SG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not found");
    SG_REQUIRE(planes == 1 || plane_stride_elems % 8 == 0, "operand plane stride must be a multiple of 8 elements");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)(planes > 1 ? plane_stride_elems : cols * rows) * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(op) failed: %d (cols=%lld rows=%lld planes=%d pstride=%lld base=%p)",
               (int)r, cols, rows, planes, plane_stride_elems, base);
    return 0;
}

// 3-D bf16 Wg[k][Cout][Cin_p], box 64 x 64 x 1
static int make_map_wg(CUtensorMap* m, const void* base, int Cin_p, int Cout, int k) {
    EncodeTiledFn fn = get_encode_fn();
    SG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[3] = {(cuuint64_t)Cin_p, (cuuint64_t)Cout, (cuuint64_t)k};
    cuuint64_t strides[2] = {(cuuint64_t)Cin_p * 2, (cuuint64_t)Cin_p * Cout * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d (Cin_p=%d Cout=%d k=%d base=%p)", (int)r, Cin_p,
               Cout, k, base);
    return 0;
}

// SMs the persistent GEMM grids may occupy.  Data parallel: NCCL's kernels need SMs of their own while backward runs; a
// persistent grid sized for all 148 SMs leaves its displaced CTAs waiting for the collective to finish, with their share
// of the tiles (static round-robin) - the whole GEMM then ends when the collective does.  sg_set_sm_limit(148 - r) keeps
// r SMs free for the collectives for as long as they are in flight (Trainer: backward only).  0 = no limit.
static int g_sm_limit = 0;

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    if (g_sm_limit > 0 && g_sm_limit < n) return g_sm_limit & ~1;
    return n;
}

// Split-K factor.  Work items = tiles x splits are dealt round-robin to `sms` persistent CTAs (CTA pairs), so the time is
// ceil(items / sms) waves of iters / splits k-blocks each.  Round 1 only split when tiles < sms; the 75-400-tile layers
// then ran 2-5 waves at 54-90 % occupancy (profiles/r1_step_profile_b64.txt: the C -> 5C k1 weight gradients, 80 tiles
// on 74 pairs, 450 TFLOP/s).  Now every shape takes the smallest factor whose last wave is >= 90 % full, as long as a
// split keeps >= 16 k-blocks (the pipeline prologue / epilogue and the red.add traffic of the partial tiles grow with it).
static int pick_splits(int tiles, int iters, int sms, bool only_underfilled = false) {
    if (iters < 32 || (only_underfilled && tiles >= sms)) return 1;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 16; ++s) {
        if (s > 1 && iters / s < 16) break;
        double waves = (double)tiles * s / sms;
        double eff = waves / ceil(waves);
        if (s > 1) eff *= 1.0 - 0.01 * s;                  // a split is not free: prefer the smaller factor on ties
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
        if (eff >= 0.90) break;
    }
    return best;
}

static void default_formats(TcParams& p) {
    p.a_fmt = p.b_fmt = c_op16_half ? 0 : 1;    // kind::f16 operand formats: 0 = fp16, 1 = bf16
    const char* e = getenv("SG_TC_FMT");     // experiment (scripts/fmt_experiment.py): "<a><b>", 0 = fp16, 1 = bf16
    if (e != nullptr && e[0] && e[1]) { p.a_fmt = e[0] - '0'; p.b_fmt = e[1] - '0'; }
}

template <int MODE>
static int launch_tc(const CUtensorMap& a, const CUtensorMap& b, TcParams p, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        SG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    int work = p.m_tiles * p.n_tiles * p.z_count * p.splits;
    int grid = work < num_sms() ? work : num_sms();
    conv_gemm_tc_kernel<MODE><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, b, p);
    return check_launch("conv_gemm_tc");
}

}  // namespace sg
#include "gemm_tc2.cuh"
namespace sg {

template <int MODE>
static int launch_tc2(const CUtensorMap& a, const CUtensorMap& b, TcParams p, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pair::conv_gemm_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             pair::SMEM_BYTES);
        SG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(pair smem=%d) failed: %s", pair::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    int work = p.m_tiles * p.n_tiles * p.z_count * p.splits;
    int pairs = num_sms() / 2;
    int grid = 2 * (work < pairs ? work : pairs);
    pair::conv_gemm_tc2_kernel<MODE><<<grid, pair::NUM_THREADS, pair::SMEM_BYTES, st>>>(a, b, p);
    return check_launch("conv_gemm_tc2");
}

// CTA-pair kernel for every GEMM with more than one 128-row tile; SG_TC_PAIR=0 forces the single-CTA kernel.
static bool use_pair(int M) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SG_TC_PAIR");
        mode = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return mode == 1 && M > 128;
}

// rowstat != NULL: ask the epilogue for per-(sample, row) GroupNorm partial sums; *stats_done tells the caller
// whether the launched kernel produced them (CTA-pair kernel without split-K) or a separate pass is needed.
int tc_fprop(const void* wg, const void* act, int act_planes, long long act_pstride, const float* bias, void* out,
             int out_bf16, int Cin, int Cin_p, int Cout, int k, int R, int accumulate, float* rowstat, int st_T, int st_Tp,
             int* stats_done, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (make_map_wg(&ma, wg, Cin_p, Cout, k)) return 1;
    if (make_map_op(&mb, act, R, Cin, act_planes, act_pstride)) return 1;
    TcParams p{};
    default_formats(p);
    const bool pair = use_pair(Cout);
    SG_REQUIRE(!out_bf16 || pair, "conv_fprop: 16-bit output needs the CTA-pair kernel (Cout > 128)");
    p.out = (float*)out; p.bias = bias; p.M = Cout; p.N = R; p.ldc = R; p.c_sz = 0;
    p.m_tiles = (int)cdiv(Cout, pair ? pair::PM : BM); p.n_tiles = (int)cdiv(R, BN); p.z_count = 1; p.group_m = 8;
    p.taps = k; p.kblocks = (int)cdiv(Cin, BK); p.pad = k / 2; p.accumulate = accumulate;
    p.b_plane0 = act_planes / 2 - k / 2; p.b_plane_step = 1; p.a_plane = 0;
    // a conv that feeds a GroupNorm keeps its statistics in the epilogue: split only when the grid would be underfilled
    p.splits = pick_splits(p.m_tiles * p.n_tiles, p.taps * p.kblocks, pair ? num_sms() / 2 : num_sms(), rowstat != nullptr);
    if (out_bf16) p.splits = 1;
    p.atomic = p.splits > 1;
    p.m_fastest = p.m_tiles <= p.n_tiles;
    p.out_bf16 = out_bf16;
    const bool fused_stats = rowstat != nullptr && pair && p.splits == 1 && !accumulate && st_Tp > 0 && R % st_Tp == 0;
    if (stats_done) *stats_done = fused_stats ? 1 : 0;
    if (fused_stats) {
        p.rowstat = rowstat; p.st_T = st_T; p.st_Tp = st_Tp; p.st_B = R / st_Tp;
        cudaMemsetAsync(rowstat, 0, sizeof(float) * 2 * (size_t)Cout * p.st_B, st);
    }
    if (p.atomic && !accumulate) cudaMemsetAsync(out, 0, sizeof(float) * (size_t)Cout * R, st);
    if (pair) return launch_tc2<MODE_FPROP>(ma, mb, p, st);
    return launch_tc<MODE_FPROP>(ma, mb, p, st);
}

// stats[b][g] = (mean, rstd) from the epilogue's per-row partial sums rowstat[b][m][2]; grid (G, B)
__global__ void __launch_bounds__(256) gn_rowstat_finalize_kernel(const float* __restrict__ rowstat, float* __restrict__ mr,
                                                                  int M, int G, double inv_n) {
    __shared__ double sh[2][32];
    const int g = blockIdx.x, b = blockIdx.y, Cg = M / G;
    const float2* rs = reinterpret_cast<const float2*>(rowstat) + (size_t)b * M + (size_t)g * Cg;
    double s = 0.0, ss = 0.0;
    for (int c = threadIdx.x; c < Cg; c += blockDim.x) {
        float2 v = rs[c];
        s += (double)v.x;
        ss += (double)v.y;
    }
    double t0 = block_sum(s, sh[0]);
    double t1 = block_sum(ss, sh[1]);
    if (threadIdx.x == 0) {
        double m = t0 * inv_n;
        double var = t1 * inv_n - m * m;
        if (var < 0.0) var = 0.0;
        mr[2 * (b * G + g)] = (float)m;
        mr[2 * (b * G + g) + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
    }
}

int tc_dgrad(const void* wg, const void* dy, int dy_planes, long long dy_pstride, void* dx, int out16, int Cin, int Cin_p,
             int Cout, int k, int R, int accumulate, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (make_map_wg(&ma, wg, Cin_p, Cout, k)) return 1;
    if (make_map_op(&mb, dy, R, Cout, dy_planes, dy_pstride)) return 1;
    TcParams p{};
    default_formats(p);
    const bool pair = use_pair(Cin);
    SG_REQUIRE(!out16 || pair, "conv_dgrad: 16-bit output needs the CTA-pair kernel (Cin > 128)");
    p.out = (float*)dx; p.bias = nullptr; p.M = Cin; p.N = R; p.ldc = R; p.c_sz = 0;
    p.m_tiles = (int)cdiv(Cin, pair ? pair::PM : BM); p.n_tiles = (int)cdiv(R, BN); p.z_count = 1; p.group_m = 8;
    p.taps = k; p.kblocks = (int)cdiv(Cout, BK); p.pad = k / 2; p.accumulate = accumulate;
    p.b_plane0 = dy_planes / 2 + k / 2; p.b_plane_step = -1; p.a_plane = 0;
    p.splits = pick_splits(p.m_tiles * p.n_tiles, p.taps * p.kblocks, pair ? num_sms() / 2 : num_sms());
    if (out16) p.splits = 1;                 // 16-bit stores cannot be split-K partial sums
    p.out_bf16 = out16;
    p.atomic = p.splits > 1;
    p.m_fastest = p.m_tiles <= p.n_tiles;
    if (p.atomic && !accumulate) cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)Cin * R, st);
    if (pair) return launch_tc2<MODE_DGRAD>(ma, mb, p, st);
    return launch_tc<MODE_DGRAD>(ma, mb, p, st);
}

int tc_wgrad(const void* dy, int dy_planes, long long dy_pstride, const void* act, int act_planes,
             long long act_pstride, float* dwg, int Cin, int Cin_p, int Cout, int k, int R, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (make_map_op(&ma, dy, R, Cout, dy_planes, dy_pstride)) return 1;
    if (make_map_op(&mb, act, R, Cin, act_planes, act_pstride)) return 1;
    TcParams p{};
    default_formats(p);
    const bool pair = use_pair(Cout);
    p.out = dwg; p.bias = nullptr; p.M = Cout; p.N = Cin_p; p.ldc = Cin_p; p.c_sz = (long long)Cout * Cin_p;
    p.m_tiles = (int)cdiv(Cout, pair ? pair::PM : BM); p.n_tiles = (int)cdiv(Cin_p, BN); p.z_count = k; p.group_m = 8;
    p.taps = 1; p.kblocks = (int)cdiv(R, BK); p.pad = k / 2; p.accumulate = 0;
    p.b_plane0 = act_planes / 2 - k / 2; p.b_plane_step = 1; p.a_plane = dy_planes / 2;
    p.splits = pick_splits(p.m_tiles * p.n_tiles * k, p.kblocks, pair ? num_sms() / 2 : num_sms());
    p.atomic = p.splits > 1;
    p.m_fastest = p.m_tiles <= p.n_tiles;
    if (p.atomic) cudaMemsetAsync(dwg, 0, sizeof(float) * (size_t)k * Cout * Cin_p, st);
    if (pair) return launch_tc2<MODE_WGRAD>(ma, mb, p, st);
    return launch_tc<MODE_WGRAD>(ma, mb, p, st);
}

// fp32 validation path (gemm_simt.cu)
int simt_fprop(const float*, const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
int simt_dgrad(const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
int simt_wgrad(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);

}  // namespace sg

using namespace sg;

extern "C" {

#define SG_CONV_CHECK(name, planes)                                                                                   \
    SG_REQUIRE(R % 8 == 0 && Cin_p % 8 == 0 && Cin_p >= Cin && (k & 1) && (planes) >= k && ((planes) & 1),            \
               name ": bad shape Cin=%d Cin_p=%d k=%d R=%d planes=%d", Cin, Cin_p, k, R, (planes))

int sg_conv_fprop(const void* wg, const void* act, int act_planes, long long act_pstride, const float* bias,
                  float* out, int Cin, int Cin_p, int Cout, int k, int R, int accumulate, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CONV_CHECK("conv_fprop", act_planes);
    if (dtype == SG_F32)
        return simt_fprop((const float*)wg, (const float*)act + (long long)(act_planes / 2) * act_pstride, bias, out, Cin,
                          Cin_p, Cout, k, R, accumulate, as_stream(stream));
    return tc_fprop(wg, act, act_planes, act_pstride, bias, out, 0, Cin, Cin_p, Cout, k, R, accumulate, nullptr, 0, 0, nullptr,
                    as_stream(stream));
}

// fprop with the output in the 16-bit operand format and no statistics (the compact reconstruction conv of the static
// configuration: its GroupNorm statistics are per column, csrc/static_ops.cu)
int sg_conv_fprop16(const void* wg, const void* act, int act_planes, long long act_pstride, const float* bias, void* out,
                    int Cin, int Cin_p, int Cout, int k, int R, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CONV_CHECK("conv_fprop16", act_planes);
    SG_REQUIRE(is_op16(dtype) && sg_conv_out16_ok(Cout), "conv_fprop16: needs a 16-bit operand mode and a CTA-pair sized layer (Cout=%d)", Cout);
    return tc_fprop(wg, act, act_planes, act_pstride, bias, out, 1, Cin, Cin_p, Cout, k, R, 0, nullptr, 0, 0, nullptr,
                    as_stream(stream));
}

int sg_conv_fprop_gn(const void* wg, const void* act, int act_planes, long long act_pstride, const float* bias, void* out,
                     int out_bf16, int Cin, int Cin_p, int Cout, int k, int B, int T, int Tp, int G, float* stats,
                     double* ws, float* rowstat, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    const int R = B * Tp;
    SG_CONV_CHECK("conv_fprop_gn", act_planes);
    SG_REQUIRE(G > 0 && Cout % G == 0 && stats != nullptr && ws != nullptr, "conv_fprop_gn: bad GroupNorm arguments");
    cudaStream_t st = as_stream(stream);
    if (dtype == SG_F32) {
        SG_REQUIRE(!out_bf16, "conv_fprop_gn: bf16 output only in bf16 mode");
        if (simt_fprop((const float*)wg, (const float*)act + (long long)(act_planes / 2) * act_pstride, bias, (float*)out, Cin,
                       Cin_p, Cout, k, R, 0, st))
            return 1;
        return sg_gn_stats((const float*)out, ws, stats, Cout, B, T, Tp, G, stream);
    }
    int done = 0;
    if (tc_fprop(wg, act, act_planes, act_pstride, bias, out, out_bf16, Cin, Cin_p, Cout, k, R, 0, rowstat, T, Tp, &done, st))
        return 1;
    if (done) {
        dim3 grid(G, B);
        gn_rowstat_finalize_kernel<<<grid, 256, 0, st>>>(rowstat, stats, Cout, G, 1.0 / ((double)(Cout / G) * T));
        return check_launch("conv_fprop_gn");
    }
    SG_REQUIRE(!out_bf16, "conv_fprop_gn: bf16 output requires the fused-statistics path");
    return sg_gn_stats((const float*)out, ws, stats, Cout, B, T, Tp, G, stream);
}

int sg_conv_dgrad(const void* wg, const void* dy, int dy_planes, long long dy_pstride, void* dx, int dx_dtype, int Cin,
                  int Cin_p, int Cout, int k, int R, int accumulate, int dtype, void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CHECK_OP16(dx_dtype);
    SG_CONV_CHECK("conv_dgrad", dy_planes);
    if (dtype == SG_F32) {
        SG_REQUIRE(dx_dtype == SG_F32, "conv_dgrad: the fp32 validation mode writes fp32 gradients");
        return simt_dgrad((const float*)wg, (const float*)dy + (long long)(dy_planes / 2) * dy_pstride, (float*)dx, Cin, Cin_p,
                          Cout, k, R, accumulate, as_stream(stream));
    }
    return tc_dgrad(wg, dy, dy_planes, dy_pstride, dx, is_op16(dx_dtype) ? 1 : 0, Cin, Cin_p, Cout, k, R, accumulate,
                    as_stream(stream));
}

int sg_set_sm_limit(int sms) {
    SG_REQUIRE(sms >= 0, "set_sm_limit: negative SM count");
    g_sm_limit = (sms > 0 && sms < 16) ? 16 : sms;
    return 0;
}

/* 1 when a GEMM with M output rows runs on the CTA-pair kernel, i.e. can store a 16-bit output (fprop: M = Cout,
 * dgrad: M = Cin) */
int sg_conv_out16_ok(int M) { return use_pair(M) ? 1 : 0; }

int sg_conv_wgrad(const void* dy, int dy_planes, long long dy_pstride, const void* act, int act_planes,
                  long long act_pstride, float* dwg, int Cin, int Cin_p, int Cout, int k, int R, int dtype,
                  void* stream) {
    SG_CHECK_OP16(dtype);
    SG_CONV_CHECK("conv_wgrad", act_planes);
    SG_REQUIRE(dy_planes & 1, "conv_wgrad: dy_planes must be odd");
    if (dtype == SG_F32)
        return simt_wgrad((const float*)dy + (long long)(dy_planes / 2) * dy_pstride,
                          (const float*)act + (long long)(act_planes / 2) * act_pstride, dwg, Cin, Cin_p, Cout, k, R,
                          as_stream(stream));
    return tc_wgrad(dy, dy_planes, dy_pstride, act, act_planes, act_pstride, dwg, Cin, Cin_p, Cout, k, R,
                    as_stream(stream));
}

}  // extern "C"
