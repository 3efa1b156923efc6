// the preprocessing scan of the reference (modules/data_preprocess.py:65-165).
//   data_scaler(): MinMaxScaler(feature_range=(-0.7, 0.7)).fit(sampled rows) -> per-node (min, max) over the selected
//   rows of the [P*T, N] field matrix; scaler.transform(chunk): X *= scale_; X += min_ (two roundings, in the dtype of
//   the data: float64 for the reference's np.zeros-built datasets, float32 otherwise); SimulGen-VAE.py:281-283 then
//   transposes to [P, N, T] and casts to float32.
// Here: one coalesced pass for the fit (rows gathered by index on the device, nodes innermost), one pass for the
// transform that can write the scaled field in place AND the float32 [P][N][T] training layout (smem-tiled transpose)
// in the same sweep.  Both are HBM-bound streaming kernels; results are bit-identical to NumPy/sklearn (min/max are
// order independent; the transform uses unfused multiply and add).
#include "common.cuh"

#include <algorithm>

namespace sg {

template <typename T>
struct Vec2 {
    T x, y;
};

template <typename T>
__device__ __forceinline__ T nan_min(T a, T b) { return fmin(a, b); }   // fmin/fmax ignore a NaN operand: np.nanmin
template <typename T>
__device__ __forceinline__ T nan_max(T a, T b) { return fmax(a, b); }
template <>
__device__ __forceinline__ float nan_min<float>(float a, float b) { return fminf(a, b); }
template <>
__device__ __forceinline__ float nan_max<float>(float a, float b) { return fmaxf(a, b); }

template <typename T>
__device__ __forceinline__ T inf_of();
template <>
__device__ __forceinline__ double inf_of<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <>
__device__ __forceinline__ float inf_of<float>() { return __int_as_float(0x7f800000); }

// grid (column strips, row chunks); block 256 threads, VEC columns per thread.  part[y][N] <- min / max over the rows
// of chunk y.  rows == nullptr: rows 0..n_rows-1.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) minmax_fit_kernel(const T* __restrict__ data, const long long* __restrict__ rows,
                                                         long long n_rows, long long N, T* __restrict__ part_min,
                                                         T* __restrict__ part_max) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= N) return;
    const long long per = (n_rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per, r1 = min(n_rows, r0 + per);
    T mn[VEC], mx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { mn[v] = inf_of<T>(); mx[v] = -inf_of<T>(); }
    constexpr int UN = 4;
    long long i = r0;
    for (; i + UN <= r1; i += UN) {
        T val[UN][VEC];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = rows ? rows[i + u] : (i + u);
            const T* p = data + r * N + col;
            if (VEC == 2) {
                Vec2<T> t = *reinterpret_cast<const Vec2<T>*>(p);
                val[u][0] = t.x;
                val[u][VEC - 1] = t.y;
            } else {
                val[u][0] = *p;
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) { mn[v] = nan_min(mn[v], val[u][v]); mx[v] = nan_max(mx[v], val[u][v]); }
    }
    for (; i < r1; ++i) {
        const long long r = rows ? rows[i] : i;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T x = data[r * N + col + v];
            mn[v] = nan_min(mn[v], x);
            mx[v] = nan_max(mx[v], x);
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        part_min[(long long)blockIdx.y * N + col + v] = mn[v];
        part_max[(long long)blockIdx.y * N + col + v] = mx[v];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_finalize_kernel(const T* __restrict__ part_min, const T* __restrict__ part_max,
                                                              int Y, long long N, T* __restrict__ out_min,
                                                              T* __restrict__ out_max, int merge) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    T mn = merge ? out_min[col] : inf_of<T>(), mx = merge ? out_max[col] : -inf_of<T>();
    for (int y = 0; y < Y; ++y) {
        mn = nan_min(mn, part_min[(long long)y * N + col]);
        mx = nan_max(mx, part_max[(long long)y * N + col]);
    }
    out_min[col] = mn;
    out_max[col] = mx;
}

__device__ __forceinline__ double mul_add_rn(double x, double s, double m) { return __dadd_rn(__dmul_rn(x, s), m); }
__device__ __forceinline__ float mul_add_rn(float x, float s, float m) { return __fadd_rn(__fmul_rn(x, s), m); }
__device__ __forceinline__ float to_f32(double v) { return __double2float_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }

// X <- X * scale + min over [R][N] (out may alias data); grid-stride over rows, threads over columns
template <typename T, int VEC>
__global__ void __launch_bounds__(256) minmax_transform_kernel(const T* data, long long R, long long N,
                                                               const T* __restrict__ scale, const T* __restrict__ minv, T* out) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= N) return;
    T s[VEC], m[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s[v] = scale[col + v]; m[v] = minv[col + v]; }
    constexpr int UN = 4;                      // rows in flight per thread
    const long long gy = gridDim.y;
    long long r = blockIdx.y;
    if (VEC == 2) {
        for (; r + (UN - 1) * gy < R; r += UN * gy) {
            Vec2<T> t[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) t[u] = *reinterpret_cast<const Vec2<T>*>(data + (r + u * gy) * N + col);
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                t[u].x = mul_add_rn(t[u].x, s[0], m[0]);
                t[u].y = mul_add_rn(t[u].y, s[VEC - 1], m[VEC - 1]);
                *reinterpret_cast<Vec2<T>*>(out + (r + u * gy) * N + col) = t[u];
            }
        }
    }
    for (; r < R; r += gy) {
        const T* p = data + r * N + col;
        T* q = out + r * N + col;
        if (VEC == 2) {
            Vec2<T> t = *reinterpret_cast<const Vec2<T>*>(p);
            t.x = mul_add_rn(t.x, s[0], m[0]);
            t.y = mul_add_rn(t.y, s[VEC - 1], m[VEC - 1]);
            *reinterpret_cast<Vec2<T>*>(q) = t;
        } else {
            *q = mul_add_rn(*p, s[0], m[0]);
        }
    }
}

// Same transform, plus the float32 [P][N][T] training layout written through a 32 x 33 shared-memory tile:
// block (32, 8); tile = 32 time steps x 32 nodes of one parameter set p; grid (node tiles, time tiles, P).
template <typename T>
__global__ void __launch_bounds__(256) minmax_transform_t_kernel(const T* data, long long N, int Tn,
                                                                 const T* __restrict__ scale, const T* __restrict__ minv,
                                                                 T* out, float* __restrict__ out_t) {
    __shared__ float tile[32][33];
    const long long n0 = (long long)blockIdx.x * 32;
    const int t0 = blockIdx.y * 32;
    const long long p = blockIdx.z;
    const long long n = n0 + threadIdx.x;
    T s = 0, m = 0;
    if (n < N) { s = scale[n]; m = minv[n]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tl = threadIdx.y + j * 8;
        const int t = t0 + tl;
        if (t < Tn && n < N) {
            const long long idx = (p * Tn + t) * N + n;
            const T v = mul_add_rn(data[idx], s, m);
            if (out != nullptr) out[idx] = v;
            tile[tl][threadIdx.x] = to_f32(v);
        }
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int nl = threadIdx.y + j * 8;
        const long long nn = n0 + nl;
        if (t < Tn && nn < N) out_t[(p * N + nn) * Tn + t] = tile[threadIdx.x][nl];
    }
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <typename T>
static int fit_impl(const T* data, const long long* rows, long long n_rows, long long N, T* ws, long long ws_elems,
                    T* out_min, T* out_max, int merge, cudaStream_t st) {
    const bool vec = (N % 2 == 0) && (reinterpret_cast<uintptr_t>(data) % (2 * sizeof(T)) == 0);
    const int per_block = 256 * (vec ? 2 : 1);
    const long long strips = cdiv(N, per_block);
    long long Y = std::max<long long>(1, std::min<long long>(cdiv(4LL * num_sms(), strips), cdiv(n_rows, 16)));
    Y = std::min<long long>(Y, ws_elems / (2 * N));
    SG_REQUIRE(Y >= 1, "minmax_fit: workspace must hold at least 2 * N elements");
    dim3 grid((unsigned)strips, (unsigned)Y);
    if (vec) minmax_fit_kernel<T, 2><<<grid, 256, 0, st>>>(data, rows, n_rows, N, ws, ws + Y * N);
    else minmax_fit_kernel<T, 1><<<grid, 256, 0, st>>>(data, rows, n_rows, N, ws, ws + Y * N);
    minmax_finalize_kernel<T><<<(unsigned)cdiv(N, 256), 256, 0, st>>>(ws, ws + Y * N, (int)Y, N, out_min, out_max, merge);
    return check_launch("minmax_fit");
}

template <typename T>
static int transform_impl(const T* data, long long R, long long N, const T* scale, const T* minv, T* out, float* out_t,
                          long long Tn, cudaStream_t st) {
    if (out_t != nullptr) {
        SG_REQUIRE(Tn > 0 && R % Tn == 0 && R / Tn <= 65535 && cdiv(Tn, 32) <= 65535, "minmax_transform: rows must be P * T with P <= 65535");
        dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(Tn, 32), (unsigned)(R / Tn)), block(32, 8);
        minmax_transform_t_kernel<T><<<grid, block, 0, st>>>(data, N, (int)Tn, scale, minv, out, out_t);
        return check_launch("minmax_transform");
    }
    SG_REQUIRE(out != nullptr, "minmax_transform: no output given");
    const bool vec = (N % 2 == 0) && (reinterpret_cast<uintptr_t>(data) % (2 * sizeof(T)) == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % (2 * sizeof(T)) == 0);
    const int per_block = 256 * (vec ? 2 : 1);
    const long long strips = cdiv(N, per_block);
    const long long Y = std::max<long long>(1, std::min<long long>(R, cdiv(8LL * num_sms(), strips)));
    dim3 grid((unsigned)strips, (unsigned)Y);
    if (vec) minmax_transform_kernel<T, 2><<<grid, 256, 0, st>>>(data, R, N, scale, minv, out);
    else minmax_transform_kernel<T, 1><<<grid, 256, 0, st>>>(data, R, N, scale, minv, out);
    return check_launch("minmax_transform");
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_minmax_fit(const void* data, int is_f64, const long long* rows, long long n_rows, long long N, void* ws,
                  long long ws_elems, void* out_min, void* out_max, int merge, void* stream) {
    SG_REQUIRE((data != nullptr || n_rows == 0) && ws != nullptr && out_min != nullptr && out_max != nullptr && n_rows >= 0 && N > 0,
               "minmax_fit: bad arguments");
    if (is_f64)
        return fit_impl<double>((const double*)data, rows, n_rows, N, (double*)ws, ws_elems, (double*)out_min, (double*)out_max,
                                merge, as_stream(stream));
    return fit_impl<float>((const float*)data, rows, n_rows, N, (float*)ws, ws_elems, (float*)out_min, (float*)out_max, merge,
                           as_stream(stream));
}

int sg_minmax_transform(const void* data, int is_f64, long long R, long long N, const void* scale, const void* minv,
                        void* out, float* out_t, long long T, void* stream) {
    if (R == 0) return 0;
    SG_REQUIRE(data != nullptr && scale != nullptr && minv != nullptr && R > 0 && N > 0, "minmax_transform: bad arguments");
    if (is_f64)
        return transform_impl<double>((const double*)data, R, N, (const double*)scale, (const double*)minv, (double*)out, out_t,
                                      T, as_stream(stream));
    return transform_impl<float>((const float*)data, R, N, (const float*)scale, (const float*)minv, (float*)out, out_t, T,
                                 as_stream(stream));
}

}  // extern "C"
