// Blocked one-sided (Hestenes) Jacobi SVD for complex128 matrices of any shape.
//
// Replaces the LAPACK zgesdd calls the reference reaches through quimb's tensor_split:
// MatrixProductState.from_dense (qmprs/primitives/mps.py:242), mps.compress
// (mps.py:451-453), gate_split_ (mps.py:928-931, 968-971) and the wasted per-sweep
// from_dense (sequential.py:443).
//
// Formulation.  The min(m,n) "short" vectors of the matrix are kept as contiguous ROWS
// of a work matrix W (W = A if m < n, W = A^T otherwise), extended on the right by an
// identity block that accumulates the rotations:  Wext = [W | I].  Rows are grouped in
// blocks of 16; a round-robin tournament pairs the blocks; for every pair of a round
//   (1) k_gram   forms the 32x32 Gram matrix of the 32 rows          (streams W once),
//   (2) k_eig    diagonalises it with a parallel two-sided Jacobi in shared memory,
//   (3) k_apply  multiplies the 32 rows of Wext by the resulting unitary (streams Wext).
// Rotations are recomputed from a freshly formed Gram matrix every visit, so rounding
// in the 32x32 solve does not accumulate and the method keeps the one-sided Jacobi
// accuracy; sweeps repeat until no Gram matrix has an off-diagonal entry above tol.
// At the end sigma_j = |row_j|, Z = rows / sigma, J = the accumulated unitary:
//   m <  n :  A = J^H Sigma Z          U = J^H,  Vh = Z
//   m >= n :  A = Z^T Sigma conj(J)    U = Z^T,  Vh = conj(J)
#include "common.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int BSZ = 16;    // rows per block
constexpr int PMAX = 32;   // rows per pair (Gram order)
constexpr int TC = 32;     // tile columns
constexpr int NT = 256;

struct Geom {
    int nv, len, nvp, nbp, single, nrows, npairs, rounds;
    long long ldw;
};

Geom make_geom(int m, int n) {
    Geom g;
    g.nv = m < n ? m : n;
    g.len = m < n ? n : m;
    if (g.nv <= PMAX) {
        g.single = 1; g.nvp = g.nv; g.nbp = 1; g.nrows = g.nv; g.npairs = 1; g.rounds = 1;
    } else {
        g.single = 0;
        int nb = (g.nv + BSZ - 1) / BSZ;
        g.nbp = nb + (nb & 1);
        g.nvp = g.nbp * BSZ;
        g.nrows = PMAX; g.npairs = g.nbp / 2; g.rounds = g.nbp - 1;
    }
    g.ldw = (long long)g.len + g.nvp;
    return g;
}

// columns held per shared-memory tile for a pair of `nrows` rows (multiple of 32, >= 32)
__host__ __device__ __forceinline__ int tile_cap(int nrows) {
    int t = ((PMAX * (TC + 1)) / nrows - 1) / 32 * 32;
    return t < 32 ? 32 : t;
}

// circle-method pairing of `n` (even) players: round r in [0,n-1), slot k in [0,n/2)
__device__ __forceinline__ void circle_pair(int r, int k, int n, int& a, int& b) {
    int n1 = n - 1;
    if (k == 0) { a = r; b = n1; }
    else { a = (r + k) % n1; b = (r - k + n1) % n1; }
    if (a > b) { int t = a; a = b; b = t; }
}

__device__ __forceinline__ int pair_row(int i, int pair, int round, int nbp, int single) {
    if (single) return i;
    int bi, bj;
    circle_pair(round, pair, nbp, bi, bj);
    return i < BSZ ? bi * BSZ + i : bj * BSZ + (i - BSZ);
}

// ---------------------------------------------------------------------------------
// (1) Gram matrices.  grid = (nchunks, npairs).  G is accumulated with atomics and is
// expected to be zero on entry (k_eig re-zeroes it after loading).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_gram(const cplx* __restrict__ W, long long ldw, int len, int chunk, int round, int nbp, int single,
       int nrows, double* __restrict__ G) {
    __shared__ cplx tile[PMAX * (TC + 1)];
    __shared__ double gacc[PMAX * PMAX * 2];
    __shared__ int rows[PMAX];
    const int tid = threadIdx.x, pair = blockIdx.y;
    const int E = nrows * nrows;
    const int tcap = tile_cap(nrows), tst = tcap + 1;   // columns per tile / row stride
    if (tid < nrows) rows[tid] = pair_row(tid, pair, round, nbp, single);
    for (int i = tid; i < E * 2; i += NT) gacc[i] = 0.0;
    const int ns = (E >= NT) ? 1 : NT / E;           // column slices per entry
    cplx acc[4];
#pragma unroll
    for (int s = 0; s < 4; s++) acc[s] = mk(0.0, 0.0);
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < len) ? c0 + chunk : len;
    __syncthreads();
    for (long long ct = c0; ct < c1; ct += tcap) {
        int tc = (int)((c1 - ct < tcap) ? (c1 - ct) : tcap);
        for (int idx = tid; idx < nrows * tcap; idx += NT) {
            int r = idx / tcap, c = idx % tcap;
            tile[r * tst + c] = (c < tc) ? W[(long long)rows[r] * ldw + ct + c] : mk(0.0, 0.0);
        }
        __syncthreads();
        if (ns == 1) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                int e = tid + s * NT;
                if (e < E) {
                    int i = e / nrows, j = e % nrows;
                    cplx a = acc[s];
                    for (int c = 0; c < tc; c++) cfmac(a, tile[i * tst + c], tile[j * tst + c]);
                    acc[s] = a;
                }
            }
        } else {
            int e = tid % E, sl = tid / E;
            if (sl < ns) {
                int i = e / nrows, j = e % nrows;
                cplx a = acc[0];
                for (int c = sl; c < tc; c += ns) cfmac(a, tile[i * tst + c], tile[j * tst + c]);
                acc[0] = a;
            }
        }
        __syncthreads();
    }
    if (ns == 1) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            int e = tid + s * NT;
            if (e < E) { gacc[2 * e] = acc[s].x; gacc[2 * e + 1] = acc[s].y; }
        }
    } else {
        int e = tid % E, sl = tid / E;
        if (sl < ns) { atomicAdd(&gacc[2 * e], acc[0].x); atomicAdd(&gacc[2 * e + 1], acc[0].y); }
    }
    __syncthreads();
    double* Gp = G + (long long)pair * PMAX * PMAX * 2;
    for (int i = tid; i < E * 2; i += NT) atomicAdd(&Gp[i], gacc[i]);
}

// ---------------------------------------------------------------------------------
// (2) Hermitian eigen-solve of each Gram matrix: parallel cyclic two-sided Jacobi.
// grid = npairs.  Writes the accumulated row transformation Q (W_new = Q W_old), the
// diagonal (squared row norms), and re-zeroes G.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_eig(double* __restrict__ G, cplx* __restrict__ Qout, int nrows, double tol2, int max_inner, int round,
      int nbp, int single, int* __restrict__ notconv, int* __restrict__ rotated, double* __restrict__ sig2) {
    __shared__ cplx g[PMAX][PMAX + 1], q[PMAX][PMAX + 1];
    __shared__ double rc[PMAX / 2], rs[PMAX / 2], rd[PMAX / 2];
    __shared__ cplx ru[PMAX / 2];
    __shared__ int rp[PMAX / 2], rq[PMAX / 2], ract[PMAX / 2];
    __shared__ int s_any, s_sweep, s_off;
    const int tid = threadIdx.x, pair = blockIdx.x;
    const int n = nrows, ne = n + (n & 1), np = ne / 2;
    double* Gp = G + (long long)pair * PMAX * PMAX * 2;
    for (int e = tid; e < n * n; e += NT) {
        int i = e / n, j = e % n;
        g[i][j] = mk(Gp[2 * e], Gp[2 * e + 1]);
        Gp[2 * e] = 0.0; Gp[2 * e + 1] = 0.0;
        q[i][j] = mk(i == j ? 1.0 : 0.0, 0.0);
    }
    if (tid == 0) { s_any = 0; s_off = 0; }
    __syncthreads();
    // is the fresh Gram matrix already diagonal to tolerance?
    {
        int off = 0;
        for (int e = tid; e < n * n; e += NT) {
            int i = e / n, j = e % n;
            if (i < j) {
                double a = g[i][i].x, b = g[j][j].x;
                if (a > 0.0 && b > 0.0 && cabs2(g[i][j]) > tol2 * a * b) off = 1;
            }
        }
        if (off) s_off = 1;
    }
    __syncthreads();
    if (s_off) {
        for (int sweep = 0; sweep < max_inner; sweep++) {
            if (tid == 0) s_sweep = 0;
            __syncthreads();
            for (int r = 0; r < ne - 1; r++) {
                if (tid < np) {
                    int p, qq;
                    if (ne == 2) { p = 0; qq = 1; } else circle_pair(r, tid, ne, p, qq);
                    int act = 0;
                    if (p < n && qq < n) {
                        double a = g[p][p].x, b = g[qq][qq].x;
                        cplx gpq = g[p][qq];
                        double mag2 = cabs2(gpq);
                        if (a > 0.0 && b > 0.0 && mag2 > tol2 * a * b) {
                            double mag = sqrt(mag2);
                            double zeta = (b - a) / (2.0 * mag);
                            double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                            double c = 1.0 / sqrt(1.0 + t * t);
                            rc[tid] = c; rs[tid] = c * t; rd[tid] = t * mag;
                            ru[tid] = mk(gpq.x / mag, gpq.y / mag);
                            act = 1;
                        }
                    }
                    rp[tid] = p; rq[tid] = qq; ract[tid] = act;
                    if (act) { s_sweep = 1; s_any = 1; }
                }
                __syncthreads();
                // row phase: rows p,q of g and q
                for (int item = tid; item < np * n; item += NT) {
                    int k = item / n, col = item % n;
                    if (!ract[k]) continue;
                    int p = rp[k], qq = rq[k];
                    double c = rc[k], s = rs[k];
                    cplx su = cscale(ru[k], s);            // s*u
                    cplx gp = g[p][col], gq = g[qq][col];
                    g[p][col] = csub(cscale(gp, c), cmul(su, gq));
                    g[qq][col] = cadd(cmul(cconj(su), gp), cscale(gq, c));
                    cplx qp = q[p][col], qv = q[qq][col];
                    q[p][col] = csub(cscale(qp, c), cmul(su, qv));
                    q[qq][col] = cadd(cmul(cconj(su), qp), cscale(qv, c));
                }
                __syncthreads();
                // column phase: columns p,q of g
                for (int item = tid; item < np * n; item += NT) {
                    int k = item / n, row = item % n;
                    if (!ract[k]) continue;
                    int p = rp[k], qq = rq[k];
                    double c = rc[k], s = rs[k];
                    cplx su = cscale(ru[k], s);
                    cplx gp = g[row][p], gq = g[row][qq];
                    g[row][p] = csub(cscale(gp, c), cmul(cconj(su), gq));
                    g[row][qq] = cadd(cmul(su, gp), cscale(gq, c));
                }
                __syncthreads();
                if (tid < np && ract[tid]) {
                    int p = rp[tid], qq = rq[tid];
                    double a = g[p][p].x, b = g[qq][qq].x;   // analytically a - t|g|, b + t|g|; keep computed, force real
                    g[p][p] = mk(a, 0.0);
                    g[qq][qq] = mk(b, 0.0);
                    g[p][qq] = mk(0.0, 0.0);
                    g[qq][p] = mk(0.0, 0.0);
                }
                __syncthreads();
            }
            if (!s_sweep) break;
            __syncthreads();
        }
    }
    cplx* Qp = Qout + (long long)pair * PMAX * PMAX;
    for (int e = tid; e < n * n; e += NT) Qp[e] = q[e / n][e % n];
    if (tid < n) sig2[pair_row(tid, pair, round, nbp, single)] = g[tid][tid].x;
    if (tid == 0) {
        rotated[pair] = s_any;
        if (s_off) atomicAdd(notconv, 1);
    }
}

// ---------------------------------------------------------------------------------
// (3) Row update  Wext[rows] <- Q * Wext[rows].  grid = (nchunks, npairs).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_apply(cplx* __restrict__ W, long long ldw, long long lenx, int chunk, int round, int nbp, int single,
        int nrows, const cplx* __restrict__ Q, const int* __restrict__ rotated) {
    const int pair = blockIdx.y;
    if (!rotated[pair]) return;
    __shared__ cplx qs[PMAX][PMAX + 1];
    __shared__ cplx tile[PMAX * (TC + 1)];
    __shared__ int rows[PMAX];
    const int tid = threadIdx.x, n = nrows;
    const int tcap = tile_cap(nrows), tst = tcap + 1;
    const cplx* Qp = Q + (long long)pair * PMAX * PMAX;
    for (int e = tid; e < n * n; e += NT) qs[e / n][e % n] = Qp[e];
    if (tid < n) rows[tid] = pair_row(tid, pair, round, nbp, single);
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < lenx) ? c0 + chunk : lenx;
    __syncthreads();
    for (long long ct = c0; ct < c1; ct += tcap) {
        int tc = (int)((c1 - ct < tcap) ? (c1 - ct) : tcap);
        for (int idx = tid; idx < n * tcap; idx += NT) {
            int r = idx / tcap, c = idx % tcap;
            if (c < tc) tile[r * tst + c] = W[(long long)rows[r] * ldw + ct + c];
        }
        __syncthreads();
        for (int idx = tid; idx < n * tcap; idx += NT) {
            int i = idx / tcap, c = idx % tcap;
            if (c < tc) {
                cplx a = mk(0.0, 0.0);
                for (int j = 0; j < n; j++) cfma(a, qs[i][j], tile[j * tst + c]);
                W[(long long)rows[i] * ldw + ct + c] = a;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// layout kernels: coalesced row gather and shared-memory tiled transpose, both with
// optional conjugation, row permutation and 1/S scaling (used to build Wext from A
// and to emit U / Vh).  These are the streaming passes of the TT-SVD.
// ---------------------------------------------------------------------------------
// out[r][c] = f(in[perm[r]][c]) * (S ? 1/S[r] : 1),  r < rows, c < cols
__global__ void k_rowcopy(cplx* __restrict__ out, long long ldo, const cplx* __restrict__ in, long long ldi,
                          const int* __restrict__ perm, const double* __restrict__ S, int conj, int rows,
                          long long cols) {
    int r = blockIdx.y;
    int src = perm ? perm[r] : r;
    double sc = 1.0;
    if (S) { double s = S[r]; sc = (s > 0.0) ? 1.0 / s : 0.0; }
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cols;
         c += (long long)gridDim.x * blockDim.x) {
        cplx v = in[(long long)src * ldi + c];
        if (conj) v.y = -v.y;
        out[(long long)r * ldo + c] = cscale(v, sc);
    }
}

// out[a][j] = f(in[perm[j]][a]) * (S ? 1/S[j] : 1),  j < nsel, a < len
__global__ void k_transpose(cplx* __restrict__ out, long long ldo, const cplx* __restrict__ in, long long ldi,
                            const int* __restrict__ perm, const double* __restrict__ S, int conj, int nsel,
                            long long len, int na) {
    __shared__ cplx tile[32][33];
    long long a0 = (long long)(blockIdx.x % na) * 32;
    int j0 = (int)(blockIdx.x / na) * 32;
    int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int jj = ty; jj < 32; jj += 8) {
        int j = j0 + jj;
        long long a = a0 + tx;
        cplx v = mk(0.0, 0.0);
        if (j < nsel && a < len) {
            int src = perm ? perm[j] : j;
            v = in[(long long)src * ldi + a];
            if (conj) v.y = -v.y;
            if (S) { double s = S[j]; v = cscale(v, (s > 0.0) ? 1.0 / s : 0.0); }
        }
        tile[jj][tx] = v;
    }
    __syncthreads();
    for (int aa = ty; aa < 32; aa += 8) {
        long long a = a0 + aa;
        int j = j0 + tx;
        if (a < len && j < nsel) out[a * ldo + j] = tile[tx][aa];
    }
}

// identity block of Wext and zero padding rows
__global__ void k_init_ext(cplx* __restrict__ W, long long ldw, int nv, int nvp, int len) {
    int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nvp; c += gridDim.x * blockDim.x)
        W[(long long)r * ldw + len + c] = mk(r == c ? 1.0 : 0.0, 0.0);
    if (r >= nv)
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < len;
             c += (long long)gridDim.x * blockDim.x)
            W[(long long)r * ldw + c] = mk(0.0, 0.0);
}

// rank by counting: S sorted descending, perm[rank] = source row
__global__ void k_sort(const double* __restrict__ sig2, int nv, double* __restrict__ S, int* __restrict__ perm) {
    extern __shared__ double s2[];
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        double v = sig2[i];
        s2[i] = (v > 0.0) ? v : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        double v = s2[j];
        int rank = 0;
        for (int k = 0; k < nv; k++) {
            double w = s2[k];
            rank += (w > v || (w == v && k < j)) ? 1 : 0;
        }
        perm[rank] = j;
        S[rank] = sqrt(v);
    }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Work {
    cplx* W; double* G; cplx* Q; int* rotated; double* sig2; int* perm; int* notconv;
    size_t total;
};

Work carve(const Geom& g, void* base) {
    Work w;
    size_t off = 0;
    char* b = (char*)base;
    w.W = (cplx*)(b + off); off += align_up((size_t)g.nvp * g.ldw * sizeof(cplx));
    w.G = (double*)(b + off); off += align_up((size_t)g.npairs * PMAX * PMAX * 2 * sizeof(double));
    w.Q = (cplx*)(b + off); off += align_up((size_t)g.npairs * PMAX * PMAX * sizeof(cplx));
    w.rotated = (int*)(b + off); off += align_up((size_t)g.npairs * sizeof(int));
    w.sig2 = (double*)(b + off); off += align_up((size_t)g.nvp * sizeof(double));
    w.perm = (int*)(b + off); off += align_up((size_t)g.nvp * sizeof(int));
    w.notconv = (int*)(b + off); off += align_up(sizeof(int));
    w.total = off;
    return w;
}

}  // namespace

extern "C" long long qm_svd_work_bytes(int m, int n) {
    Geom g = make_geom(m, n);
    Work w = carve(g, nullptr);
    return (long long)w.total;
}

// A (m x n, row-major, lda) is not modified.  U: m x k (ldu), S: k, Vh: k x n (ldvh), k = min(m,n).
// U or Vh may be NULL.  info_host (optional, host int[2]) receives {sweeps, converged}.
extern "C" int qm_svd(int m, int n, const void* A_, long long lda, void* U_, long long ldu, void* S_, void* Vh_,
                      long long ldvh, void* work, long long work_bytes, double tol, int max_sweeps,
                      int* info_host, void* stream_) {
    if (m <= 0 || n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    Geom g = make_geom(m, n);
    if (g.nv > 6144) return -2;   // k_sort shared memory bound (48 KB)
    Work w = carve(g, work);
    if ((long long)w.total > work_bytes) return -1;
    const cplx* A = (const cplx*)A_;
    double* S = (double*)S_;

    // --- build Wext = [W | I] ---
    if (m < n) {
        dim3 grid(ceil_div(n, 256) > 4096 ? 4096 : ceil_div(n, 256), m);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_rowcopy<<<grid, 256, 0, st>>>(w.W, g.ldw, A, lda, nullptr, nullptr, 0, m, n));
    } else {
        // W[j][a] = A[a][j]: transpose of the m x n input
        // k_transpose maps in[perm[j]][a] -> out[a][j]; here "in" = A, out = W (n x m):
        // W[c][r] = A[r][c]  =>  nsel = m (rows of A), len = n (cols of A)
        int na = ceil_div(n, 32);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)na * ceil_div(m, 32)), dim3(32, 8), 0, st>>>(
            w.W, g.ldw, A, lda, nullptr, nullptr, 0, m, n, na));
    }
    QM_CHECK_LAUNCH();
    {
        dim3 grid(ceil_div(g.nvp > 256 ? g.nvp : 256, 256), g.nvp);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_init_ext<<<grid, 256, 0, st>>>(w.W, g.ldw, g.nv, g.nvp, g.len));
        QM_CHECK_LAUNCH();
    }
    QM_CUDA(cudaMemsetAsync(w.G, 0, (size_t)g.npairs * PMAX * PMAX * 2 * sizeof(double), st));
    QM_CUDA(cudaMemsetAsync(w.sig2, 0, (size_t)g.nvp * sizeof(double), st));

    // --- sweeps ---
    const long long lenx = g.ldw;
    // chunking: aim for >= ~2 waves of CTAs but at least 256 columns per CTA
    auto pick_chunk = [&](long long cols) {
        long long want = (long long)(600 / (g.npairs > 0 ? g.npairs : 1));
        if (want < 1) want = 1;
        long long chunk = (cols + want - 1) / want;
        if (chunk < 256) chunk = 256;
        chunk = (chunk + TC - 1) / TC * TC;
        return chunk;
    };
    const long long chunk_g = pick_chunk(g.len), chunk_a = pick_chunk(lenx);
    const int ncg = ceil_div(g.len, chunk_g), nca = ceil_div(lenx, chunk_a);
    const double tol2 = tol * tol;
    int sweeps = 0, converged = 0;
    for (; sweeps < max_sweeps;) {
        QM_CUDA(cudaMemsetAsync(w.notconv, 0, sizeof(int), st));
        for (int r = 0; r < g.rounds; r++) {
            QM_LAUNCH(QM_CLS_SVD_GRAM, st, k_gram<<<dim3(ncg, g.npairs), NT, 0, st>>>(w.W, g.ldw, g.len, (int)chunk_g, r, g.nbp, g.single,
                                                        g.nrows, w.G));
            QM_LAUNCH(QM_CLS_SVD_EIG, st, k_eig<<<g.npairs, NT, 0, st>>>(w.G, w.Q, g.nrows, tol2, 12, r, g.nbp, g.single, w.notconv,
                                           w.rotated, w.sig2));
            QM_LAUNCH(QM_CLS_SVD_APPLY, st, k_apply<<<dim3(nca, g.npairs), NT, 0, st>>>(w.W, g.ldw, lenx, (int)chunk_a, r, g.nbp, g.single,
                                                         g.nrows, w.Q, w.rotated));
        }
        // complex MAC = 8 flops: Gram nrows^2 x len, update nrows^2 x lenx, per pair and round
        qm_prof_work(QM_CLS_SVD_GRAM, 8.0 * g.nrows * g.nrows * (double)g.len * g.npairs * g.rounds);
        qm_prof_work(QM_CLS_SVD_APPLY, 8.0 * g.nrows * g.nrows * (double)lenx * g.npairs * g.rounds);
        QM_CHECK_LAUNCH();
        sweeps++;
        int h = 0;
        QM_CUDA(cudaMemcpyAsync(&h, w.notconv, sizeof(int), cudaMemcpyDeviceToHost, st));
        QM_CUDA(cudaStreamSynchronize(st));
        if (h == 0) { converged = 1; break; }
    }
    if (info_host) { info_host[0] = sweeps; info_host[1] = converged; }

    // --- sort, emit ---
    QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_sort<<<1, 1024, (size_t)g.nv * sizeof(double), st>>>(w.sig2, g.nv, S, w.perm));
    QM_CHECK_LAUNCH();
    const int k = g.nv;
    cplx* U = (cplx*)U_;
    cplx* Vh = (cplx*)Vh_;
    if (m < n) {
        // U[a][j] = conj(J[perm[j]][a]);  Vh[j][c] = W[perm[j]][c] / S[j]
        if (U)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(m, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                U, ldu, w.W + g.len, g.ldw, w.perm, nullptr, 1, k, m, ceil_div(m, 32)));
        if (Vh) {
            dim3 grid(ceil_div(n, 256) > 4096 ? 4096 : ceil_div(n, 256), k);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_rowcopy<<<grid, 256, 0, st>>>(Vh, ldvh, w.W, g.ldw, w.perm, S, 0, k, n));
        }
    } else {
        // U[a][j] = W[perm[j]][a] / S[j];  Vh[j][c] = conj(J[perm[j]][c])
        if (U)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(m, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                U, ldu, w.W, g.ldw, w.perm, S, 0, k, m, ceil_div(m, 32)));
        if (Vh) {
            dim3 grid(ceil_div(n, 256), k);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_rowcopy<<<grid, 256, 0, st>>>(Vh, ldvh, w.W + g.len, g.ldw, w.perm, nullptr, 1, k, n));
        }
    }
    QM_CHECK_LAUNCH();
    return 0;
}
