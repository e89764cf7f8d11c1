// Complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64), row-major NN.
//
//   C[m x n] = alpha * op(A)[m x k] * B[k x n] + beta * C,   op(A) = A or A^H
//
// This is the dense contraction under rows A2/A4/A6/A7 of SURVEY.md section 8
// (quimb tensordot / gate_split contraction reached from qmprs/primitives/mps.py:928-931,
// 968-971, 451-453, 270).  tcgen05 has no FP64 kind, so the tensor path for complex128
// is warp-level DMMA with operands staged through shared memory as split re/im planes:
//   Cre += Are*Bre + (-Aim)*Bim ;  Cim += Are*Bim + Aim*Bre      (4 DMMAs per 8x8x4 tile)
#include <cstdlib>
#include <map>
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
           long long strideA, long long strideB, long long strideC, int ktotal) {
    __shared__ double sAr[BM * AS], sAi[BM * AS];
    // split-K launch: slice z covers k columns starting at z * k of ktotal (the last slice may be shorter)
    if (ktotal > 0 && (long long)(blockIdx.z + 1) * k > ktotal) k = ktotal - blockIdx.z * k;
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


// ---------------------------------------------------------------------------------
// Large-shape path: 64 x 128 x 16 tiles, 3-stage ring filled by the TMA engine (cp.async.bulk row
// copies completing on mbarriers), one producer warp + 8 DMMA warps (2 x 4, warp tile 32 x 32).
// Operands stay interleaved (re, im) in shared memory: a lane's 16-byte fragment load carries both
// parts, and the padded row strides (A: 20, B: 130 complex; A^H: 66) make those loads conflict free.
// One DMMA occupies an SM sub-partition's FP64 tensor pipe for ~16 cycles, so 64 DMMAs per 8 fragment
// loads keep the pipe busy; what the old kernel lacked is the overlap of tile loads with the math.
// ---------------------------------------------------------------------------------
constexpr int TM = 64, TN = 128, TK = 16, NSTAGE = 3;
constexpr int TAS = TK + 4;        // A tile [m][k] row stride (complex)
constexpr int TATS = TM + 2;       // A^H: tile stored [k][m]
constexpr int TBS = TN + 2;        // B tile [k][n] row stride
constexpr int A_ELEMS = TM * TAS > TK * TATS ? TM * TAS : TK * TATS;
constexpr int STAGE_ELEMS = A_ELEMS + TK * TBS;
constexpr size_t TMA_SMEM = (size_t)NSTAGE * STAGE_ELEMS * sizeof(cplx) + 2 * NSTAGE * sizeof(unsigned long long);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_row(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

template <bool TA>
__global__ void __launch_bounds__(288, 1)
k_zgemm_tma(int m, int n, int k, cplx alpha, const cplx* __restrict__ A, long long lda,
            const cplx* __restrict__ B, long long ldb, cplx beta, cplx* __restrict__ C, long long ldc,
            long long strideA, long long strideB, long long strideC, int ktotal) {
    if (ktotal > 0 && (long long)(blockIdx.z + 1) * k > ktotal) k = ktotal - blockIdx.z * k;
    extern __shared__ __align__(16) unsigned char gsm[];
    cplx* tiles = (cplx*)gsm;
    unsigned long long* full = (unsigned long long*)(tiles + (size_t)NSTAGE * STAGE_ELEMS);
    unsigned long long* empty = full + NSTAGE;

    A += (long long)blockIdx.z * strideA;
    B += (long long)blockIdx.z * strideB;
    C += (long long)blockIdx.z * strideC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int mv = m - m0 < TM ? m - m0 : TM, nv = n - n0 < TN ? n - n0 : TN;   // valid rows / cols of this tile
    const int KT = (k + TK - 1) / TK;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ---- producer: one bulk copy per tile row ----
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % NSTAGE;
            if (kt >= NSTAGE) mbar_wait(&empty[s], ((kt / NSTAGE) - 1) & 1);
            cplx* sA = tiles + (size_t)s * STAGE_ELEMS;
            cplx* sB = sA + A_ELEMS;
            const int k0 = kt * TK;
            const int kv = k - k0 < TK ? k - k0 : TK;
            if (kv < TK) {
                // k tail: the rest of the tile must be finite zeros (0 * stale NaN would poison valid outputs)
                if (TA) { for (int e = lane; e < (TK - kv) * TATS; e += 32) sA[kv * TATS + e] = mk(0.0, 0.0); }
                else { for (int e = lane; e < TM * (TK - kv); e += 32) sA[(e / (TK - kv)) * TAS + kv + e % (TK - kv)] = mk(0.0, 0.0); }
                for (int e = lane; e < (TK - kv) * TBS; e += 32) sB[kv * TBS + e] = mk(0.0, 0.0);
                __syncwarp();
            }
            const unsigned bytes = (unsigned)((TA ? kv * mv : mv * kv) + kv * nv) * 16u;
            if (lane == 0) mbar_expect_tx(&full[s], bytes);
            __syncwarp();
            if (TA) {
                for (int r = lane; r < kv; r += 32)
                    tma_row(sA + r * TATS, A + (long long)(k0 + r) * lda + m0, (unsigned)mv * 16u, &full[s]);
            } else {
                for (int r = lane; r < mv; r += 32)
                    tma_row(sA + r * TAS, A + (long long)(m0 + r) * lda + k0, (unsigned)kv * 16u, &full[s]);
            }
            for (int r = lane; r < kv; r += 32)
                tma_row(sB + r * TBS, B + (long long)(k0 + r) * ldb + n0, (unsigned)nv * 16u, &full[s]);
        }
        return;
    }

    // ---- consumers ----
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    double cr[4][4][2], ci[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

    for (int kt = 0; kt < KT; kt++) {
        const int s = kt % NSTAGE;
        mbar_wait(&full[s], (kt / NSTAGE) & 1);
        const cplx* sA = tiles + (size_t)s * STAGE_ELEMS;
        const cplx* sB = sA + A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            cplx a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int row = wm * 32 + i * 8 + g;
                if (TA) { a[i] = sA[(kk + t) * TATS + row]; a[i].y = -a[i].y; }
                else a[i] = sA[row * TAS + kk + t];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sB[(kk + t) * TBS + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    dmma884(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
                    dmma884(cr[i][j][0], cr[i][j][1], -a[i].y, b[j].y);
                    dmma884(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
                    dmma884(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
    const bool unit_alpha = (alpha.x == 1.0 && alpha.y == 0.0);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = n0 + wn * 32 + j * 8 + 2 * t + e;
                if (col >= n) continue;
                cplx v = mk(cr[i][j][e], ci[i][j][e]);
                if (!unit_alpha) v = cmul(alpha, v);
                cplx* dst = C + (long long)row * ldc + col;
                if (has_beta) v = cadd(v, cmul(beta, *dst));
                *dst = v;
            }
        }
    }
}


// ---------------------------------------------------------------------------------
// Skinny shapes of the chi=2 truncation: products with <= 8 right-hand columns and
// 4x4 / 2x2 Hermitian forms T^H M.  These are bandwidth problems (A is read once); the tiled kernels
// above would run them on a handful of CTAs with a barrier per 16 k.
// ---------------------------------------------------------------------------------
// C[m x n] = alpha * A[m x k] B[k x n] + beta C, n <= 8: one warp per row, lanes stride over k.
template <int NC>
__global__ void __launch_bounds__(256)
k_zgemm_fewcols(int m, int n, int k, cplx alpha, const cplx* __restrict__ A, long long lda,
                const cplx* __restrict__ B, long long ldb, cplx beta, cplx* __restrict__ C, long long ldc) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const cplx* a = A + row * lda;
    cplx acc[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) acc[j] = mk(0.0, 0.0);
    for (int kk = lane; kk < k; kk += 32) {
        const cplx av = a[kk];
        const cplx* b = B + (long long)kk * ldb;
#pragma unroll
        for (int j = 0; j < NC; j++)
            if (j < n) cfma(acc[j], av, b[j]);
    }
#pragma unroll
    for (int j = 0; j < NC; j++) { acc[j].x = warp_sum(acc[j].x); acc[j].y = warp_sum(acc[j].y); }
    if (lane == 0) {
        const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
        for (int j = 0; j < NC; j++)
            if (j < n) {
                cplx v = cmul(alpha, acc[j]);
                cplx* dst = C + row * ldc + j;
                if (has_beta) v = cadd(v, cmul(beta, *dst));
                *dst = v;
            }
    }
}

// C[m x n] = alpha * A^H B + beta C with A (k x m), B (k x n), m, n <= 4: one CTA, threads stride over k,
// fixed-order block reduction.
__global__ void __launch_bounds__(1024)
k_zgemm_ta_tiny(int m, int n, int k, cplx alpha, const cplx* __restrict__ A, long long lda,
                const cplx* __restrict__ B, long long ldb, cplx beta, cplx* __restrict__ C, long long ldc) {
    __shared__ double red[32][33];
    cplx acc[16];
#pragma unroll
    for (int e = 0; e < 16; e++) acc[e] = mk(0.0, 0.0);
    for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {
        cplx av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) av[i] = i < m ? A[(long long)kk * lda + i] : mk(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < 4; j++) bv[j] = j < n ? B[(long long)kk * ldb + j] : mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) ccfma(acc[i * 4 + j], av[i], bv[j]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const double x = warp_sum(acc[e].x), y = warp_sum(acc[e].y);
        if (lane == 0) { red[warp][2 * e] = x; red[warp][2 * e + 1] = y; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double ssum = 0.0;
        for (int w = 0; w < nw; w++) ssum += red[w][threadIdx.x];
        red[0][threadIdx.x] = ssum;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int i = threadIdx.x >> 2, j = threadIdx.x & 3;
        if (i < m && j < n) {
            cplx v = cmul(alpha, mk(red[0][2 * threadIdx.x], red[0][2 * threadIdx.x + 1]));
            cplx* dst = C + (long long)i * ldc + j;
            if (beta.x != 0.0 || beta.y != 0.0) v = cadd(v, cmul(beta, *dst));
            *dst = v;
        }
    }
}


// C = alpha * sum_s P[s] + beta * C over the split-K partial slabs (fixed order: deterministic)
__global__ void k_splitk_reduce(cplx* __restrict__ C, long long ldc, const cplx* __restrict__ P, int splits, int m, int n,
                                cplx alpha, cplx beta) {
    const long long total = (long long)m * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        cplx acc = mk(0.0, 0.0);
        for (int sidx = 0; sidx < splits; sidx++) acc = cadd(acc, P[(long long)sidx * total + e]);
        cplx v = cmul(alpha, acc);
        cplx* dst = C + (e / n) * ldc + (e % n);
        if (beta.x != 0.0 || beta.y != 0.0) v = cadd(v, cmul(beta, *dst));
        *dst = v;
    }
}

// grow-only split-K workspace per stream (the pipeline runs large states on one stream)
struct SplitWs { void* p; size_t bytes; };
std::map<cudaStream_t, SplitWs> g_split_ws;
void* split_ws(cudaStream_t st, size_t bytes) {
    SplitWs& w = g_split_ws[st];
    if (w.bytes < bytes) {
        if (w.p) { cudaStreamSynchronize(st); cudaFree(w.p); }
        if (cudaMalloc(&w.p, bytes) != cudaSuccess) { w.p = nullptr; w.bytes = 0; return nullptr; }
        w.bytes = bytes;
    }
    return w.p;
}

}  // namespace

extern "C" int qm_zgemm(int m, int n, int k, double alpha_re, double alpha_im, const void* A, long long lda,
                        const void* B, long long ldb, double beta_re, double beta_im, void* C, long long ldc,
                        int batch, long long strideA, long long strideB, long long strideC, int trans_a,
                        void* stream) {
    if (m <= 0 || n <= 0 || batch <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // large shapes: TMA-fed pipelined kernel (needs 16-byte aligned rows, true for complex128 matrices)
    static int use_tma = -1;
    if (use_tma < 0) {
        const char* e = getenv("QM_GEMM_TMA");
        use_tma = e ? atoi(e) : 1;
        if (cudaFuncSetAttribute(k_zgemm_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_zgemm_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess)
            return -4;
    }
    // few output tiles but a long k (left environments, Gram products): split k over extra CTAs, partial slabs
    // summed in fixed order
    {
        const bool tma_ok = use_tma && m >= TM && n >= TN;
        const long long tiles = tma_ok ? (long long)ceil_div(m, TM) * ceil_div(n, TN)
                                       : (long long)ceil_div(m, BM) * ceil_div(n, BN);
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        // k from 256 up, slices of >= 64 (round 2: k >= 512, slices >= 128 -- a 256^3 product then ran on 16 CTAs, 6.6x
        // behind cuBLAS; QM_GEMM_SPLITK_MINK=512 restores that rule for A/B runs)
        static const int splitk_mink = getenv("QM_GEMM_SPLITK_MINK") ? atoi(getenv("QM_GEMM_SPLITK_MINK")) : 256;
        const int min_slice = splitk_mink >= 512 ? 128 : 64;
        if (batch == 1 && tiles < 64 && k >= splitk_mink && (long long)m * n >= 4096 && cap == cudaStreamCaptureStatusNone) {
            int splits = (int)(148 / tiles);               // one wave: tiles x splits <= 148 CTAs (512^3: 5 x 32 = 160 CTAs measured slower than 4 x 32)
            if (splits > 16) splits = 16;
            if (splits > k / min_slice) splits = k / min_slice;
            int kc = ((k + splits - 1) / splits + 15) / 16 * 16;
            splits = (k + kc - 1) / kc;
            if (splits >= 2) {
                cplx* P = (cplx*)split_ws(st, (size_t)splits * m * n * sizeof(cplx));
                if (!P) return -5;
                const long long sA = trans_a ? (long long)kc * lda : (long long)kc, sB = (long long)kc * ldb;
                const long long sC = (long long)m * n;
                const cplx one = mk(1.0, 0.0), zero = mk(0.0, 0.0);
                if (tma_ok) {
                    dim3 grid(ceil_div(m, TM), ceil_div(n, TN), splits);
                    if (trans_a)
                        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_tma<true><<<grid, 288, TMA_SMEM, st>>>(
                            m, n, kc, one, (const cplx*)A, lda, (const cplx*)B, ldb, zero, P, n, sA, sB, sC, k)));
                    else
                        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_tma<false><<<grid, 288, TMA_SMEM, st>>>(
                            m, n, kc, one, (const cplx*)A, lda, (const cplx*)B, ldb, zero, P, n, sA, sB, sC, k)));
                } else {
                    dim3 grid(ceil_div(m, BM), ceil_div(n, BN), splits);
                    if (trans_a)
                        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<true><<<grid, 256, 0, st>>>(
                            m, n, kc, one, (const cplx*)A, lda, (const cplx*)B, ldb, zero, P, n, sA, sB, sC, k)));
                    else
                        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<false><<<grid, 256, 0, st>>>(
                            m, n, kc, one, (const cplx*)A, lda, (const cplx*)B, ldb, zero, P, n, sA, sB, sC, k)));
                }
                const int nb = ceil_div((long long)m * n, 256) > 148 * 8 ? 148 * 8 : ceil_div((long long)m * n, 256);
                QM_LAUNCH(QM_CLS_GEMM, st, (k_splitk_reduce<<<nb, 256, 0, st>>>((cplx*)C, ldc, P, splits, m, n,
                                                                               mk(alpha_re, alpha_im), mk(beta_re, beta_im))));
                qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k);
                QM_CHECK_LAUNCH();
                return 0;
            }
        }
    }
    if (use_tma && m >= TM && n >= TN && k >= 2 * TK && (long long)ceil_div(m, TM) * ceil_div(n, TN) * batch >= 16) {
        dim3 grid(ceil_div(m, TM), ceil_div(n, TN), batch);
        if (grid.y > 65535u) return -3;
        if (trans_a) {
            QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_tma<true><<<grid, 288, TMA_SMEM, st>>>(
                m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
                (cplx*)C, ldc, strideA, strideB, strideC, 0)));
        } else {
            QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_tma<false><<<grid, 288, TMA_SMEM, st>>>(
                m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
                (cplx*)C, ldc, strideA, strideB, strideC, 0)));
        }
        qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k * batch);
        QM_CHECK_LAUNCH();
        return 0;
    }
    if (batch == 1 && !trans_a && n <= 8 && k >= 64 && m >= 8) {
        const cplx al = mk(alpha_re, alpha_im), be = mk(beta_re, beta_im);
        const int nb = ceil_div(m, 8);
        if (n <= 2)
            QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_fewcols<2><<<nb, 256, 0, st>>>(m, n, k, al, (const cplx*)A, lda, (const cplx*)B, ldb, be, (cplx*)C, ldc)));
        else if (n <= 4)
            QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_fewcols<4><<<nb, 256, 0, st>>>(m, n, k, al, (const cplx*)A, lda, (const cplx*)B, ldb, be, (cplx*)C, ldc)));
        else
            QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_fewcols<8><<<nb, 256, 0, st>>>(m, n, k, al, (const cplx*)A, lda, (const cplx*)B, ldb, be, (cplx*)C, ldc)));
        qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k);
        QM_CHECK_LAUNCH();
        return 0;
    }
    if (batch == 1 && trans_a && m <= 4 && n <= 4 && k >= 64) {
        const int nt = k >= 1024 ? 1024 : ((k + 31) / 32) * 32;
        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm_ta_tiny<<<1, nt, 0, st>>>(m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda,
                                                                     (const cplx*)B, ldb, mk(beta_re, beta_im), (cplx*)C, ldc)));
        qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k);
        QM_CHECK_LAUNCH();
        return 0;
    }
    dim3 grid(ceil_div(m, BM), ceil_div(n, BN), batch);
    if (grid.y > 65535u) return -3;
    if (trans_a) {
        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<true><<<grid, 256, 0, st>>>(
            m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
            (cplx*)C, ldc, strideA, strideB, strideC, 0)));
    } else {
        QM_LAUNCH(QM_CLS_GEMM, st, (k_zgemm<false><<<grid, 256, 0, st>>>(
            m, n, k, mk(alpha_re, alpha_im), (const cplx*)A, lda, (const cplx*)B, ldb, mk(beta_re, beta_im),
            (cplx*)C, ldc, strideA, strideB, strideC, 0)));
    }
    qm_prof_work(QM_CLS_GEMM, 8.0 * m * n * (double)k * batch);
    QM_CHECK_LAUNCH();
    return 0;
}
