// Complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64), row-major NN.
//
//   C[m x n] = alpha * op(A)[m x k] * B[k x n] + beta * C,   op(A) = A or A^H
//
// This is the dense contraction under rows A2/A4/A6/A7 of SURVEY.md section 8
// (quimb tensordot / gate_split contraction reached from qmprs/primitives/mps.py:928-931,
// 968-971, 451-453, 270).  tcgen05 has no FP64 kind, so the tensor path for complex128
// is warp-level DMMA with operands staged through shared memory as split re/im planes:
//   Cre += Are*Bre + (-Aim)*Bim ;  Cim += Are*Bim + Aim*Bre      (4 DMMAs per 8x8x4 tile)
#include "common.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int AS = BK + 4;   // padded row stride (doubles) of the A planes: conflict-free fragment loads
constexpr int BS = BN + 4;   // padded row stride of the B planes

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// 256 threads = 8 warps laid out 4 (m) x 2 (n); warp tile 16 x 32 = 2 x 4 DMMA tiles.
// TA: op(A) = A^H, A stored (k x m).
template <bool TA>
__global__ void __launch_bounds__(256)
k_zgemm(int m, int n, int k, cplx alpha, const cplx* __restrict__ A, long long lda,
           const cplx* __restrict__ B, long long ldb, cplx beta, cplx* __restrict__ C, long long ldc,
           long long strideA, long long strideB, long long strideC) {
    __shared__ double sAr[BM * AS], sAi[BM * AS];
    __shared__ double sBr[BK * BS], sBi[BK * BS];

    A += (long long)blockIdx.z * strideA;
    B += (long long)blockIdx.z * strideB;
    C += (long long)blockIdx.z * strideC;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // x covers M (can be 2^23 rows in to_dense)
    const int g = lane >> 2, t = lane & 3;

    double cr[2][4][2], ci[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

    for (int k0 = 0; k0 < k; k0 += BK) {
        // A tile: BM x BK
        for (int idx = tid; idx < BM * BK; idx += 256) {
            int r, c;
            if (TA) { c = idx / BM; r = idx % BM; } else { r = idx / BK; c = idx % BK; }
            cplx v = mk(0.0, 0.0);
            if (m0 + r < m && k0 + c < k) {
                if (TA) v = cconj(A[(long long)(k0 + c) * lda + (m0 + r)]);
                else v = A[(long long)(m0 + r) * lda + (k0 + c)];
            }
            sAr[r * AS + c] = v.x;
            sAi[r * AS + c] = v.y;
        }
        // B tile: BK x BN
        for (int idx = tid; idx < BK * BN; idx += 256) {
            int r = idx / BN, c = idx % BN;
            cplx v = mk(0.0, 0.0);
            if (k0 + r < k && n0 + c < n) v = B[(long long)(k0 + r) * ldb + (n0 + c)];
            sBr[r * BS + c] = v.x;
            sBi[r * BS + c] = v.y;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double ar[2], ai[2], nai[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                int row = wm * 16 + i * 8 + g;
                ar[i] = sAr[row * AS + kk + t];
                ai[i] = sAi[row * AS + kk + t];
                nai[i] = -ai[i];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int col = wn * 32 + j * 8 + g;
                double br = sBr[(kk + t) * BS + col];
                double bi = sBi[(kk + t) * BS + col];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    dmma884(cr[i][j][0], cr[i][j][1], ar[i], br);
                    dmma884(cr[i][j][0], cr[i][j][1], nai[i], bi);
                    dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi);
                    dmma884(ci[i][j][0], ci[i][j][1], ai[i], br);
                }
            }
        }
        __syncthreads();
    }

    const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int row = m0 + wm * 16 + i * 8 + g;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int col = n0 + wn * 32 + j * 8 + 2 * t + e;
                if (col >= n) continue;
                cplx v = cmul(alpha, mk(cr[i][j][e], ci[i][j][e]));
                cplx* dst = C + (long long)row * ldc + col;
                if (has_beta) v = cadd(v, cmul(beta, *dst));
                *dst = v;
            }
        }
    }
}

}  // namespace

extern "C" int qm_zgemm(int m, int n, int k, double alpha_re, double alpha_im, const void* A, long long lda,
                        const void* B, long long ldb, double beta_re, double beta_im, void* C, long long ldc,
                        int batch, long long strideA, long long strideB, long long strideC, int trans_a,
                        void* stream) {
    if (m <= 0 || n <= 0 || batch <= 0) return 0;
    dim3 grid(ceil_div(m, BM), ceil_div(n, BN), batch);
    if (grid.y > 65535u) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    if (trans_a) {
        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<true><<<grid, 256, 0, st>>>(
            m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
            (cplx*)C, ldc, strideA, strideB, strideC)));
    } else {
        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<false><<<grid, 256, 0, st>>>(
            m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
            (cplx*)C, ldc, strideA, strideB, strideC)));
    }
    qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k * batch);
    QM_CHECK_LAUNCH();
    return 0;
}
