// fp32 validation-mode convolutions (SIMT, no tensor cores).  Same operand layouts and the same
// three entry points as the tcgen05 path (gemm_tc.cu); used when dtype == SG_F32 so that every
// other kernel and the whole wiring can be checked against the fp32 oracle to ~1e-5 relative L2
// (BASELINE.json north_star "fp32 validation mode").  Not a performance path.
#include "common.cuh"

namespace sg {

constexpr int TM = 64, TN = 64, TK = 16;

// C[z?][m][n] (+)= sum_z sum_k A[z*a_sz + m*a_sm + k*a_sk] * B[k*b_sk + n*b_sn + shift_z]
// where the operand index that runs along the CR "r" axis (n when r_is_n, else k) is only valid
// while 0 <= r + shift_z < R (zero padding).  shift_z = shift0 + z*shift_step.
// sum_over_z: all z accumulate into one C (fprop/dgrad); otherwise z = blockIdx.z picks C + z*c_sz (wgrad).
struct SimtArgs {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    int M, N, K, Z;
    long long a_sz, a_sm, a_sk, b_sk, b_sn, c_sz, c_sm;
    int shift0, shift_step, r_is_n, R, sum_over_z, accumulate;
};

__global__ void __launch_bounds__(256) simt_conv_kernel(SimtArgs p) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    int z_lo = p.sum_over_z ? 0 : blockIdx.z, z_hi = p.sum_over_z ? p.Z : blockIdx.z + 1;
    for (int z = z_lo; z < z_hi; ++z) {
        int shift = p.shift0 + z * p.shift_step;
        const float* Az = p.A + (long long)z * p.a_sz;
        for (int k0 = 0; k0 < p.K; k0 += TK) {
            for (int e = threadIdx.x; e < TK * TM; e += 256) {
                int kk, mm;
                if (p.a_sk == 1) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
                int m = m0 + mm, k = k0 + kk;
                float v = 0.f;
                if (m < p.M && k < p.K) v = Az[(long long)m * p.a_sm + (long long)k * p.a_sk];
                As[kk][mm] = v;
            }
            for (int e = threadIdx.x; e < TK * TN; e += 256) {
                int kk, nn;
                if (p.b_sn == 1) { nn = e % TN; kk = e / TN; } else { kk = e % TK; nn = e / TK; }
                int n = n0 + nn, k = k0 + kk;
                float v = 0.f;
                if (n < p.N && k < p.K) {
                    int r = (p.r_is_n ? n : k) + shift;
                    if (r >= 0 && r < p.R) v = p.B[(long long)k * p.b_sk + (long long)n * p.b_sn + shift];
                }
                Bs[kk][nn] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
            }
            __syncthreads();
        }
    }
    float* Cz = p.C + (p.sum_over_z ? 0 : (long long)blockIdx.z * p.c_sz);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        float bv = p.bias ? p.bias[m] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            long long ci = (long long)m * p.c_sm + n;
            float v = acc[i][j] + bv;
            Cz[ci] = p.accumulate ? Cz[ci] + v : v;
        }
    }
}

static int launch_simt(const SimtArgs& p, cudaStream_t st) {
    dim3 grid((unsigned)cdiv(p.N, TN), (unsigned)cdiv(p.M, TM), p.sum_over_z ? 1 : p.Z);
    simt_conv_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("simt_conv");
}

int simt_fprop(const float* wg, const float* act, const float* bias, float* out, int Cin, int Cin_p, int Cout, int k,
               int R, int accumulate, cudaStream_t st) {
    SimtArgs p{};
    p.A = wg; p.B = act; p.C = out; p.bias = bias;
    p.M = Cout; p.N = R; p.K = Cin; p.Z = k;
    p.a_sz = (long long)Cout * Cin_p; p.a_sm = Cin_p; p.a_sk = 1;
    p.b_sk = R; p.b_sn = 1; p.c_sz = 0; p.c_sm = R;
    p.shift0 = -(k / 2); p.shift_step = 1; p.r_is_n = 1; p.R = R; p.sum_over_z = 1; p.accumulate = accumulate;
    return launch_simt(p, st);
}

int simt_dgrad(const float* wg, const float* dy, float* dx, int Cin, int Cin_p, int Cout, int k, int R, int accumulate,
               cudaStream_t st) {
    SimtArgs p{};
    p.A = wg; p.B = dy; p.C = dx; p.bias = nullptr;
    p.M = Cin; p.N = R; p.K = Cout; p.Z = k;
    p.a_sz = (long long)Cout * Cin_p; p.a_sm = 1; p.a_sk = Cin_p;
    p.b_sk = R; p.b_sn = 1; p.c_sz = 0; p.c_sm = R;
    p.shift0 = k / 2; p.shift_step = -1; p.r_is_n = 1; p.R = R; p.sum_over_z = 1; p.accumulate = accumulate;
    return launch_simt(p, st);
}

int simt_wgrad(const float* dy, const float* act, float* dwg, int Cin, int Cin_p, int Cout, int k, int R,
               cudaStream_t st) {
    SimtArgs p{};
    p.A = dy; p.B = act; p.C = dwg; p.bias = nullptr;
    p.M = Cout; p.N = Cin; p.K = R; p.Z = k;
    p.a_sz = 0; p.a_sm = R; p.a_sk = 1;
    p.b_sk = 1; p.b_sn = R; p.c_sz = (long long)Cout * Cin_p; p.c_sm = Cin_p;
    p.shift0 = -(k / 2); p.shift_step = 1; p.r_is_n = 0; p.R = R; p.sum_over_z = 0; p.accumulate = 0;
    return launch_simt(p, st);
}

}  // namespace sg
