// Single-CTA Jacobi SVD of small complex128 matrices, the whole problem in shared memory; grid = batch.
//
// Replaces, for registers of <= 12-13 qubits, the same LAPACK zgesdd calls as svd.cu (quimb tensor_split behind
// qmprs/primitives/mps.py:242, :451-453, :968-971).  In BASELINE config 5 (12 qubits, chi=64) every matrix of the
// path has min(m,n) <= 64 and m*n <= 4096: the multi-kernel solver of svd.cu spends ~50-120 launches on each of
// the ~120 SVDs of a state (5000 of its 6500 CUDA-graph nodes); here an SVD is ONE node that occupies one SM, so
// the states of a batch overlap on the 148 SMs instead of queueing behind each other's launches.
//
// Method: one-sided (Hestenes) Jacobi on the min(m,n) short vectors kept as rows of W (W = A or A^T), optionally
// extended by an identity block that accumulates the rotations (as svd.cu).  Row pairs of a round-robin round are
// rotated concurrently, a pair by a segment of 16 or 32 lanes: dot products by shuffles inside the segment, the
// rotation formula of svd.cu, one block barrier per round; sweeps until no pair has |<x,y>|^2 > tol^2 |x|^2 |y|^2.
// Deterministic (fixed shuffle trees, no atomics on data).
#include "common.cuh"
#include "qmprs_b200.h"
#include <cstdlib>

namespace {

constexpr int NTS = 256;                                  // threads per CTA (max; see qm_svd_small)
constexpr int MAXNV = 64;
constexpr int EPL = 4;                                    // row elements a lane keeps in registers across a rotation
constexpr size_t SMEM_CAP = 200 * 1024;

__device__ __forceinline__ void circle_pair_s(int r, int k, int n, int& a, int& b) {
    const int n1 = n - 1;
    if (k == 0) { a = r; b = n1; }
    else { a = (r + k) % n1; b = (r - k + n1) % n1; }
    if (a > b) { const int t = a; a = b; b = t; }
}

struct Dims { int nv, len, ext, ldw; };
__host__ __device__ __forceinline__ Dims small_dims(int m, int n, int backmult) {
    Dims d;
    d.nv = m < n ? m : n;
    d.len = m < n ? n : m;
    d.ext = backmult ? 0 : d.nv;
    d.ldw = d.len + d.ext;
    if ((d.ldw & 1) == 0) d.ldw += 1;                     // odd row stride: conflict-free 16-byte column accesses
    return d;
}
__host__ __device__ __forceinline__ size_t small_smem(const Dims& d) {
    return (size_t)d.nv * d.ldw * sizeof(cplx) + (size_t)MAXNV * (sizeof(double) + sizeof(int)) + 64;
}

// 2 CTAs per SM (64 registers): co-resident problems of a batch overlap their shuffle / FP64 / barrier chains; measured
// on config 5: layer-extraction phase 713 -> 611-638 ms with 3 per SM and no register caching -- the solver is bound by
// shared-memory traffic (a 64 x 64 round moves 3 x 64 KB), hence the caching below
template <bool CACHE>
__global__ void __launch_bounds__(NTS, CACHE ? 2 : 3)
k_svd_small(int m, int n, const cplx* __restrict__ A_, long long lda, long long sA, cplx* __restrict__ U_, long long ldu,
            long long sU, double* __restrict__ S_, long long sS, cplx* __restrict__ Vh_, long long ldvh, long long sVh,
            double tol2, int max_sweeps, int backmult, int* __restrict__ mismatch) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    const Dims d = small_dims(m, n, backmult);
    const int nv = d.nv, len = d.len, ext = d.ext, ldw = d.ldw, lenx = d.len + d.ext;
    cplx* W = (cplx*)ss_smem;
    double* sig = (double*)(W + (size_t)nv * ldw);
    int* perm = (int*)(sig + MAXNV);
    const int tid = threadIdx.x, lane = tid & 31, nts = blockDim.x;     // 256 or 512 threads (host choice)
    const cplx* A = A_ + (long long)blockIdx.x * sA;
    // ---- load W = A (m < n) or A^T, identity extension ----
    if (m < n) {
        for (int idx = tid; idx < m * n; idx += nts) {
            const int i = idx / n, c = idx % n;
            W[i * ldw + c] = A[(long long)i * lda + c];
        }
    } else {
        for (int idx = tid; idx < m * n; idx += nts) {
            const int a = idx / n, j = idx % n;
            W[j * ldw + a] = A[(long long)a * lda + j];
        }
    }
    for (int idx = tid; idx < nv * ext; idx += nts) {
        const int i = idx / ext, c = idx % ext;
        W[i * ldw + len + c] = mk(i == c ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    // ---- sweeps ----
    const int ne = nv + (nv & 1), npairs = ne / 2;
    int tpp = 32;                                          // lanes per pair (power of two, a pair never spans warps)
    while (tpp > 1 && npairs * tpp > nts) tpp >>= 1;
    const int k = tid / tpp, j = tid % tpp;
    const bool has_pair = k < npairs;
    const bool cached = CACHE && len <= EPL * tpp;
    int converged = nv < 2 ? 1 : 0;
    for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
        int any = 0;
        for (int r = 0; r < ne - 1; r++) {
            int p = 0, q = 0;
            bool valid = false;
            if (has_pair) {
                if (ne == 2) { p = 0; q = 1; } else circle_pair_s(r, k, ne, p, q);
                valid = q < nv;                            // odd nv: the dummy partner sits out
            }
            cplx* wp = W + p * ldw;
            cplx* wq = W + q * ldw;
            double a = 0.0, b = 0.0;
            cplx g = mk(0.0, 0.0);
            // a lane's elements of the two rows stay in registers from the dot products to the rotation when they
            // are at most EPL each (always in config 5): the solver is bound by shared-memory traffic, 3 -> 2 passes
            cplx xr[EPL], yr[EPL];
            if (valid) {
                if (cached) {
#pragma unroll
                    for (int e = 0; e < EPL; e++) {
                        const int c = j + e * tpp;
                        xr[e] = c < len ? wp[c] : mk(0.0, 0.0);
                        yr[e] = c < len ? wq[c] : mk(0.0, 0.0);
                    }
#pragma unroll
                    for (int e = 0; e < EPL; e++) {      // same order of accumulation as the loop below
                        a += cabs2(xr[e]);
                        b += cabs2(yr[e]);
                        cfmac(g, xr[e], yr[e]);
                    }
                } else {
                    for (int c = j; c < len; c += tpp) {
                        const cplx x = wp[c], y = wq[c];
                        a += cabs2(x);
                        b += cabs2(y);
                        cfmac(g, x, y);                        // x conj(y)
                    }
                }
            }
            for (int off = tpp >> 1; off > 0; off >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off);
                b += __shfl_xor_sync(0xffffffffu, b, off);
                g.x += __shfl_xor_sync(0xffffffffu, g.x, off);
                g.y += __shfl_xor_sync(0xffffffffu, g.y, off);
            }
            const double mag2 = cabs2(g);
            if (valid && a > 0.0 && b > 0.0 && mag2 > tol2 * a * b) {
                // rotation of svd.cu (overflow-free, no 1/|g|): row p <- c x + o y ; row q <- -conj(o) x + c y
                const double dd = 0.5 * (b - a);
                const double hh = fma(dd, dd, mag2);
                const double den = fabs(dd) + hh * rsqrt(hh);
                const double R = rsqrt(fma(den, den, mag2));
                const double s = copysign(R, dd);
                const double cc = den * R;
                const cplx o = mk(-s * g.x, -s * g.y);
                int c0 = j;
                if (cached) {
#pragma unroll
                    for (int e = 0; e < EPL; e++) {
                        const int c = j + e * tpp;
                        if (c < len) {
                            wp[c] = cadd(cscale(xr[e], cc), cmul(o, yr[e]));
                            wq[c] = csub(cscale(yr[e], cc), cmul(cconj(o), xr[e]));
                        }
                    }
                    c0 = j + ((len - j + tpp - 1) / tpp) * tpp;      // first column >= len of this lane
                }
                for (int c = c0; c < lenx; c += tpp) {
                    const cplx x = wp[c], y = wq[c];
                    wp[c] = cadd(cscale(x, cc), cmul(o, y));
                    wq[c] = csub(cscale(y, cc), cmul(cconj(o), x));
                }
                // a sweep whose rotations all had |cos| <= 1e-8 leaves cosines of ~1e-16 n (quadratic convergence of the
                // cyclic Jacobi iteration): it ends the iteration, no verification sweep (40 % of a sweep's cost)
                if (mag2 > 1e-16 * a * b) any = 1;
            }
            __syncthreads();
        }
        if (!__syncthreads_or(any)) converged = 1;
    }
    if (!converged && mismatch && tid == 0) mismatch[0] = 1;
    // ---- singular values, sorted descending by counting rank ----
    for (int i = tid >> 5; i < nv; i += nts / 32) {
        double s2 = 0.0;
        for (int c = lane; c < len; c += 32) s2 += cabs2(W[i * ldw + c]);
        s2 = warp_sum(s2);
        if (lane == 0) sig[i] = s2;
    }
    __syncthreads();
    if (tid < nv) {
        const double v = sig[tid];
        int rank = 0;
        for (int t = 0; t < nv; t++) {
            const double w = sig[t];
            rank += (w > v || (w == v && t < tid)) ? 1 : 0;
        }
        perm[rank] = tid;
    }
    __syncthreads();
    double* S = S_ + (long long)blockIdx.x * sS;
    if (tid < nv) S[tid] = sqrt(sig[perm[tid]]);
    cplx* U = U_ ? U_ + (long long)blockIdx.x * sU : nullptr;
    cplx* Vh = Vh_ ? Vh_ + (long long)blockIdx.x * sVh : nullptr;
    const int kk = nv;
    if (m < n) {
        // rows of W are sigma_j z_j:  Vh[j][c] = W[perm[j]][c] / sigma_j ;  U = J^H  or  A Z^H Sigma^-1
        if (Vh)
            for (int idx = tid; idx < kk * n; idx += nts) {
                const int jj = idx / n, c = idx % n, src = perm[jj];
                const double s2 = sig[src];
                Vh[(long long)jj * ldvh + c] = cscale(W[src * ldw + c], s2 > 0.0 ? rsqrt(s2) : 0.0);
            }
        if (U)
            for (int idx = tid; idx < m * kk; idx += nts) {
                const int a = idx / kk, jj = idx % kk, src = perm[jj];
                cplx v;
                if (ext) v = cconj(W[src * ldw + len + a]);
                else {
                    v = mk(0.0, 0.0);
                    const cplx* arow = A + (long long)a * lda;
                    for (int c = 0; c < n; c++) cfmac(v, arow[c], W[src * ldw + c]);      // A conj(W_j)
                    const double s2 = sig[src];
                    v = cscale(v, s2 > 0.0 ? 1.0 / s2 : 0.0);
                }
                U[(long long)a * ldu + jj] = v;
            }
    } else {
        // rows of W are sigma_j u_j^T:  U[a][j] = W[perm[j]][a] / sigma_j ;  Vh = conj(J)  or  Sigma^-1 U^H A
        if (U)
            for (int idx = tid; idx < m * kk; idx += nts) {
                const int a = idx / kk, jj = idx % kk, src = perm[jj];
                const double s2 = sig[src];
                U[(long long)a * ldu + jj] = cscale(W[src * ldw + a], s2 > 0.0 ? rsqrt(s2) : 0.0);
            }
        if (Vh)
            for (int idx = tid; idx < kk * n; idx += nts) {
                const int jj = idx / n, c = idx % n, src = perm[jj];
                cplx v;
                if (ext) v = cconj(W[src * ldw + len + c]);
                else {
                    v = mk(0.0, 0.0);
                    for (int a = 0; a < m; a++) ccfma(v, W[src * ldw + a], A[(long long)a * lda + c]);   // conj(W_j) A
                    const double s2 = sig[src];
                    v = cscale(v, s2 > 0.0 ? 1.0 / s2 : 0.0);
                }
                Vh[(long long)jj * ldvh + c] = v;
            }
    }
}

}  // namespace

// 1 if an m x n problem runs in the single-CTA solver (min(m,n) <= 64 and the work matrix fits shared memory)
extern "C" int qm_svd_small_fits(int m, int n, int flags) {
    if (m < 1 || n < 1) return 0;
    const Dims d = small_dims(m, n, flags & QM_SVD_BACKMULT);
    return d.nv <= MAXNV && small_smem(d) <= SMEM_CAP;
}

// `batch` independent thin SVDs A_b = U_b diag(S_b) Vh_b of the same shape, one CTA each, no workspace, no host
// synchronisation (CUDA-graph capturable).  Strides in elements; U / Vh may be NULL; S sorted descending.
// mismatch (optional int[1]) is set to 1 if a problem has not converged after max_sweeps sweeps.
// Returns -3 if the shape does not fit (qm_svd_small_fits).
extern "C" int qm_svd_small(int m, int n, const void* A, long long lda, long long strideA, void* U, long long ldu,
                            long long strideU, void* S, long long strideS, void* Vh, long long ldvh, long long strideVh,
                            double tol, int max_sweeps, int flags, int batch, void* mismatch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (batch < 1) return 0;
    const int backmult = (flags & QM_SVD_BACKMULT) ? 1 : 0;
    if (!qm_svd_small_fits(m, n, flags)) return -3;
    const Dims d = small_dims(m, n, backmult);
    const size_t smem = small_smem(d);
    // QM_SVD_SMALL_CACHE=1: row elements cached in registers across a rotation, 64 registers, 2 CTAs per SM;
    // 0 (default): 40 registers, 3 CTAs per SM (measured on config 5: profiles/bench_r02_c5_*.json)
    static const int cache = getenv("QM_SVD_SMALL_CACHE") ? atoi(getenv("QM_SVD_SMALL_CACHE")) : 0;
    static bool attr_set = false;
    if (!attr_set) {
        // largest shared-memory carve-out for every launch: the batch path interleaves these CTAs with other small
        // kernels on the same SMs, and CTAs that need different carve-outs cannot share an SM
        QM_CUDA(cudaFuncSetAttribute(k_svd_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
        QM_CUDA(cudaFuncSetAttribute(k_svd_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
        QM_CUDA(cudaFuncSetAttribute(k_svd_small<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        QM_CUDA(cudaFuncSetAttribute(k_svd_small<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    // threads per CTA: 256 -- a row pair of a 64-row problem gets 8 lanes instead of 16 (one shuffle stage less, the
    // rotation formula evaluated by half as many lanes) and three CTAs of a batch share an SM.  Measured on config 5
    // (layer-extraction phase of 4096 states): 512 threads 637 ms, 256: 499 ms, 128: 608 ms.  QM_SVD_SMALL_THREADS overrides.
    static const int thr_env = getenv("QM_SVD_SMALL_THREADS") ? atoi(getenv("QM_SVD_SMALL_THREADS")) : 0;
    const int nthreads = (thr_env >= 64 && thr_env <= NTS && thr_env % 32 == 0) ? thr_env : 256;
    if (cache)
        QM_LAUNCH(QM_CLS_SVD_EIG, st, k_svd_small<true><<<batch, nthreads, smem, st>>>(
            m, n, (const cplx*)A, lda, strideA, (cplx*)U, ldu, strideU, (double*)S, strideS, (cplx*)Vh, ldvh, strideVh,
            tol * tol, max_sweeps, backmult, (int*)mismatch));
    else
        QM_LAUNCH(QM_CLS_SVD_EIG, st, k_svd_small<false><<<batch, nthreads, smem, st>>>(
            m, n, (const cplx*)A, lda, strideA, (cplx*)U, ldu, strideU, (double*)S, strideS, (cplx*)Vh, ldvh, strideVh,
            tol * tol, max_sweeps, backmult, (int*)mismatch));
    QM_CHECK_LAUNCH();
    return 0;
}
