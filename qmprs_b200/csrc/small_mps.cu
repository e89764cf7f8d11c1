// Fused MPS bookkeeping kernels for small registers (bonds <= 64: BASELINE config 5, 12 qubits / chi = 64).
//
// With every matrix of such a register inside one SM, a state is no longer bound by arithmetic but by the NUMBER of
// kernels it needs: the states of a batch are replayed from CUDA graphs and the GPU front end retires ~2.5 M graph
// kernel nodes per second (measured, profiles/bench_r02_c5_*.json).  Each kernel here replaces a fixed group of
// launches of the generic path with identical arithmetic order inside every output element:
//   qm_split_absorb   qm_trim + qm_expect_ints + 2 x qm_scale_copy after a gate-split / TT-SVD split
//                     (quimb _trim_and_renorm_svd_result + absorb, behind mps.py:242, :451-453, :968-971)
//   qm_theta_small    qm_zgemm + qm_theta_gate: two-site contraction with the 4x4 gate        (mps.py:968-971)
//   qm_chi2_env       2 x qm_zgemm: left environment L_i = sum_p B_i^H L_{i-1} B_i of the chi=2 truncation (mps.py:881)
//   qm_chi2_bond      4 x qm_zgemm + qm_chi2_select: one bond of that truncation
//   qm_zero_overlap   N-1 x qm_zgemm + qm_expect_not_close: <0..0|psi> as a product of the p=0 slices (mps.py:1020-1039)
#include "common.cuh"
#include "chi2_select.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int NT = 256;

// element strides between consecutive states of a batch, one per pointer argument (in argument order)
template <int K> struct BatchStrides { long long s[K]; };

// ---------------------------------------------------------------------------------
// split + absorb.  Every CTA recomputes the rank (k <= a few thousand values), then copies its share.
// mode 0: 'rel' cutoff (s_j > cutoff s_0), max_bond, singular values absorbed to the LEFT (left = U S, right = Vh);
// mode 1: 'rsum2' cutoff with Frobenius renormalisation f, sqrt(s f) absorbed on BOTH sides.
// expect: the rank the static pipeline assumed (output shapes); mismatch[0] = 1 if the data disagree.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_split_absorb(const cplx* __restrict__ U, long long ldu, const double* __restrict__ S, const cplx* __restrict__ Vh,
               long long ldvh, int m, int n, int k, double cutoff, int mode, int max_bond, int expect,
               cplx* __restrict__ left, cplx* __restrict__ right, int* __restrict__ mismatch, BatchStrides<5> bs) {
    // state blockIdx.y of a batch of same-shape problems (strides in elements; single problem: grid.y = 1)
    U += blockIdx.y * bs.s[0]; S += blockIdx.y * bs.s[1]; Vh += blockIdx.y * bs.s[2];
    left += blockIdx.y * bs.s[3]; right += blockIdx.y * bs.s[4]; mismatch += blockIdx.y;
    __shared__ double red[33];
    __shared__ int s_rank;
    __shared__ double s_f;
    const int tid = threadIdx.x;
    double part = 0.0;
    for (int i = tid; i < k; i += NT) part += S[i] * S[i];
    const double tot = block_sum(part, red);
    if (tid == 0) {
        int r;
        if (mode == 0) {
            const double thr = cutoff * S[0];
            r = 0;
            for (int i = 0; i < k; i++) r += (S[i] > thr) ? 1 : 0;
        } else {
            const double target = cutoff * tot;
            double ssum = 0.0;
            r = k;
            for (int i = k - 1; i >= 0; i--) {
                ssum += S[i] * S[i];
                if (ssum > target) break;
                r--;
            }
        }
        if (r < 1) r = 1;
        if (max_bond > 0 && r > max_bond) r = max_bond;
        double f = 1.0;
        if (mode == 1 && r < k) {
            double keep = 0.0, lose = 0.0;
            for (int i = 0; i < r; i++) keep += S[i] * S[i];
            for (int i = r; i < k; i++) lose += S[i] * S[i];
            f = sqrt((keep + lose) / keep);
        }
        s_rank = r;
        s_f = f;
        if (r != expect && blockIdx.x == 0) mismatch[0] = 1;
    }
    __syncthreads();
    const double f = s_f;
    const int r = expect;                                  // output shapes are the assumed ones
    const long long nl = (long long)m * r, nr = (long long)r * n;
    for (long long idx = (long long)blockIdx.x * NT + tid; idx < nl + nr; idx += (long long)gridDim.x * NT) {
        if (idx < nl) {
            const int a = (int)(idx / r), j = (int)(idx % r);
            double w = S[j];
            if (mode == 1) w = sqrt(w * f);
            left[idx] = cscale(U[(long long)a * ldu + j], w);
        } else {
            const long long e = idx - nl;
            const int j = (int)(e / n), c = (int)(e % n);
            const double w = mode == 1 ? sqrt(S[j] * f) : 1.0;
            right[e] = cscale(Vh[(long long)j * ldvh + c], w);
        }
    }
}

// ---------------------------------------------------------------------------------
// theta[(l,oi),(oj,r)] = sum M[(oi,oj),(pi,pj)] sum_b A[l,pi,b] A2[b,pj,r],  M = G or G^H.
// One thread per (l, r): the four (pi, pj) contractions over b, then the 4x4 mix.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_theta_small(const cplx* __restrict__ A, const cplx* __restrict__ A2, int l, int b, int r, const cplx* __restrict__ G,
              int dagger, cplx* __restrict__ X, BatchStrides<4> bs) {
    A += blockIdx.y * bs.s[0]; A2 += blockIdx.y * bs.s[1]; G += blockIdx.y * bs.s[2]; X += blockIdx.y * bs.s[3];
    __shared__ cplx M[16];
    if (threadIdx.x < 16) {
        const int a = threadIdx.x / 4, c = threadIdx.x % 4;
        M[threadIdx.x] = dagger ? cconj(G[c * 4 + a]) : G[a * 4 + c];
    }
    __syncthreads();
    const long long idx = (long long)blockIdx.x * NT + threadIdx.x;
    if (idx >= (long long)l * r) return;
    const int li = (int)(idx / r), ri = (int)(idx % r);
    cplx x[4] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};       // (pi, pj)
    const cplx* a0 = A + (long long)(li * 2 + 0) * b;
    const cplx* a1 = A + (long long)(li * 2 + 1) * b;
    for (int k = 0; k < b; k++) {
        const cplx u0 = a0[k], u1 = a1[k];
        const cplx v0 = A2[((long long)k * 2 + 0) * r + ri], v1 = A2[((long long)k * 2 + 1) * r + ri];
        cfma(x[0], u0, v0); cfma(x[1], u0, v1); cfma(x[2], u1, v0); cfma(x[3], u1, v1);
    }
    const long long ld = 2LL * r;
    cplx* base = X + (long long)(2 * li) * ld + ri;
    cplx y[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        cplx s = mk(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 4; c++) cfma(s, M[a * 4 + c], x[c]);
        y[a] = s;
    }
    base[0] = y[0]; base[r] = y[1]; base[ld] = y[2]; base[ld + r] = y[3];
}

// ---------------------------------------------------------------------------------
// Left environment of the chi=2 truncation:  L_i[c][c'] = sum_{a,p,x} conj(B[a,p,c]) Lprev[a][x] B[x,p,c']
// (Lprev = identity for the first site).  One CTA; Y[p][a][c'] = sum_x Lprev[a][x] B[x,p,c'] staged in shared memory.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_chi2_env(const cplx* __restrict__ Lprev, const cplx* __restrict__ B, int l, int r, cplx* __restrict__ Lout,
           BatchStrides<3> bs) {
    if (Lprev) Lprev += blockIdx.x * bs.s[0];
    B += blockIdx.x * bs.s[1]; Lout += blockIdx.x * bs.s[2];
    extern __shared__ __align__(16) unsigned char env_smem[];
    cplx* Y = (cplx*)env_smem;                             // [l*2][r]  (row index a*2 + p, as B)
    const int tid = threadIdx.x;
    for (int idx = tid; idx < l * 2 * r; idx += 512) {
        const int row = idx / r, c = idx % r, a = row >> 1, p = row & 1;
        cplx s;
        if (Lprev) {
            s = mk(0.0, 0.0);
            for (int x = 0; x < l; x++) cfma(s, Lprev[a * l + x], B[((long long)x * 2 + p) * r + c]);
        } else {
            s = B[(long long)row * r + c];
        }
        Y[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += 512) {
        const int c = idx / r, c2 = idx % r;
        cplx s = mk(0.0, 0.0);
        for (int row = 0; row < 2 * l; row++) ccfma(s, B[(long long)row * r + c], Y[row * r + c2]);
        Lout[idx] = s;
    }
}

// ---------------------------------------------------------------------------------
// One bond of the chi=2 truncation (fast path of host.chi2_layer):
//   M = L T (b x 4) ;  H = T^H M (4 x 4) ;  select (rank <= 2, canonical phases) -> C_i, Vsel ;
//   W = T Vsel (b x 2) ;  T_out = Bprev W  ((l0*2) x 2, read as (l0, 4)).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_chi2_bond(const cplx* __restrict__ L, int b, const cplx* __restrict__ T, const cplx* __restrict__ Bprev, int l0,
            double cutoff, double tie, double amb_rel, cplx* __restrict__ Csite, int* __restrict__ bond,
            int* __restrict__ ambiguous, cplx* __restrict__ Tout, BatchStrides<7> bs) {
    L += blockIdx.x * bs.s[0]; T += blockIdx.x * bs.s[1]; Bprev += blockIdx.x * bs.s[2]; Csite += blockIdx.x * bs.s[3];
    bond += blockIdx.x * bs.s[4]; ambiguous += blockIdx.x * bs.s[5]; Tout += blockIdx.x * bs.s[6];
    __shared__ cplx Ts[64 * 4], Ms[64 * 4], Ws[64 * 2], H[16], Vsel[8];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < b * 4; idx += NT) Ts[idx] = T[idx];
    __syncthreads();
    for (int idx = tid; idx < b * 4; idx += NT) {
        const int a = idx >> 2, c = idx & 3;
        cplx s = mk(0.0, 0.0);
        for (int x = 0; x < b; x++) cfma(s, L[a * b + x], Ts[x * 4 + c]);
        Ms[idx] = s;
    }
    __syncthreads();
    if (tid < 16) {
        const int c = tid >> 2, c2 = tid & 3;
        cplx s = mk(0.0, 0.0);
        for (int a = 0; a < b; a++) ccfma(s, Ts[a * 4 + c], Ms[a * 4 + c2]);
        H[tid] = s;
    }
    __syncthreads();
    if (tid == 0) chi2_select_dev(nullptr, H, 4, cutoff, tie, Csite, Vsel, bond, 2, amb_rel, ambiguous);
    __syncthreads();
    for (int idx = tid; idx < b * 2; idx += NT) {
        const int a = idx >> 1, j = idx & 1;
        cplx s = mk(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 4; c++) cfma(s, Ts[a * 4 + c], Vsel[c * 2 + j]);
        Ws[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < l0 * 2 * 2; idx += NT) {
        const int row = idx >> 1, j = idx & 1;
        cplx s = mk(0.0, 0.0);
        for (int a = 0; a < b; a++) cfma(s, Bprev[(long long)row * b + a], Ws[a * 2 + j]);
        Tout[idx] = s;
    }
}

// ---------------------------------------------------------------------------------
// v <- v B_i[:,0,:] over the sites; out = v (re, im); the early-break test |f - 1| <= tol must NOT fire
// (sequential.py:390) when tol >= 0, else mismatch.
// ---------------------------------------------------------------------------------
struct SiteList { const cplx* p[32]; int l[32]; int r[32]; long long s[32]; };     // s: batch stride of each site tensor

__global__ void __launch_bounds__(NT)
k_zero_overlap(SiteList sites, int N, double tol, cplx* __restrict__ out, int* __restrict__ mismatch) {
    __shared__ cplx va[1024], vb[1024];
    const int tid = threadIdx.x;
    cplx* v = va;
    cplx* w = vb;
    if (tid == 0) v[0] = mk(1.0, 0.0);
    __syncthreads();
    out += blockIdx.x;
    mismatch += blockIdx.x;
    for (int i = 0; i < N; i++) {
        const cplx* B = sites.p[i] + blockIdx.x * sites.s[i];
        const int l = sites.l[i], r = sites.r[i];
        for (int c = tid; c < r; c += NT) {
            cplx s = mk(0.0, 0.0);
            for (int a = 0; a < l; a++) cfma(s, v[a], B[(long long)a * 2 * r + c]);       // B[a, 0, c]
            w[c] = s;
        }
        __syncthreads();
        cplx* t = v; v = w; w = t;
    }
    if (tid == 0) {
        out[0] = v[0];
        if (tol >= 0.0) {
            const double dr = v[0].x - 1.0, di = v[0].y;
            if (sqrt(dr * dr + di * di) <= tol) mismatch[0] = 1;
        }
    }
}

}  // namespace

namespace {
template <int K> BatchStrides<K> strides_from(const long long* strides) {
    BatchStrides<K> bs;
    for (int i = 0; i < K; i++) bs.s[i] = strides ? strides[i] : 0;
    return bs;
}
}  // namespace

// The *_batch entries run `batch` same-shape problems in ONE launch (the lock-step lanes of graphs.py: W states advance
// together, so that a CUDA-graph replay has one node per step instead of W).  `strides`: HOST array with the element
// stride between consecutive problems of every pointer argument, in argument order (the mismatch / flag arguments are
// int vectors indexed by the problem).  The single-problem entries are batch = 1.
extern "C" int qm_split_absorb_batch(const void* U, long long ldu, const void* S, const void* Vh, long long ldvh, int m,
                                     int n, int k, double cutoff, int mode, int max_bond, int expect_rank, void* left,
                                     void* right, void* mismatch, int batch, const long long* strides, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (m < 1 || n < 1 || k < 1 || expect_rank < 1 || expect_rank > k || batch < 1 || batch > 65535) return -2;
    const long long work = (long long)expect_rank * (m + n);
    long long g = (work + NT - 1) / NT;
    if (g > 148 * 4) g = 148 * 4;
    QM_LAUNCH(QM_CLS_SMALL, st, k_split_absorb<<<dim3((unsigned)g, (unsigned)batch), NT, 0, st>>>(
        (const cplx*)U, ldu, (const double*)S, (const cplx*)Vh, ldvh, m, n, k, cutoff, mode, max_bond, expect_rank,
        (cplx*)left, (cplx*)right, (int*)mismatch, strides_from<5>(strides)));
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_split_absorb(const void* U, long long ldu, const void* S, const void* Vh, long long ldvh, int m, int n,
                               int k, double cutoff, int mode, int max_bond, int expect_rank, void* left, void* right,
                               void* mismatch, void* stream) {
    return qm_split_absorb_batch(U, ldu, S, Vh, ldvh, m, n, k, cutoff, mode, max_bond, expect_rank, left, right, mismatch,
                                 1, nullptr, stream);
}

extern "C" int qm_theta_small_batch(const void* A, const void* A2, int l, int b, int r, const void* G, int dagger,
                                    void* X, int batch, const long long* strides, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (batch < 1 || batch > 65535) return -2;
    QM_LAUNCH(QM_CLS_GEMM, st, k_theta_small<<<dim3((unsigned)ceil_div((long long)l * r, NT), (unsigned)batch), NT, 0, st>>>(
        (const cplx*)A, (const cplx*)A2, l, b, r, (const cplx*)G, dagger, (cplx*)X, strides_from<4>(strides)));
    qm_prof_work(QM_CLS_GEMM, 8.0 * (2.0 * l) * b * (2.0 * r) * batch);
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_theta_small(const void* A, const void* A2, int l, int b, int r, const void* G, int dagger, void* X,
                              void* stream) {
    return qm_theta_small_batch(A, A2, l, b, r, G, dagger, X, 1, nullptr, stream);
}

extern "C" int qm_chi2_env_batch(const void* Lprev, const void* B, int l, int r, void* Lout, int batch,
                                 const long long* strides, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)l * 2 * r * sizeof(cplx);
    if (smem > 200 * 1024) return -3;
    if (batch < 1) return -2;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        QM_CUDA(cudaFuncSetAttribute(k_chi2_env, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = 200 * 1024;
    }
    QM_LAUNCH(QM_CLS_GEMM, st, k_chi2_env<<<batch, 512, smem, st>>>((const cplx*)Lprev, (const cplx*)B, l, r, (cplx*)Lout,
                                                                   strides_from<3>(strides)));
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_chi2_env(const void* Lprev, const void* B, int l, int r, void* Lout, void* stream) {
    return qm_chi2_env_batch(Lprev, B, l, r, Lout, 1, nullptr, stream);
}

extern "C" int qm_chi2_bond_batch(const void* L, int b, const void* T, const void* Bprev, int l0, double cutoff,
                                  double tie, double ambiguous_rel, void* Csite, void* bond, void* ambiguous, void* Tout,
                                  int batch, const long long* strides, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (b < 1 || b > 64) return -3;
    if (batch < 1) return -2;
    QM_LAUNCH(QM_CLS_SMALL, st, k_chi2_bond<<<batch, NT, 0, st>>>((const cplx*)L, b, (const cplx*)T, (const cplx*)Bprev, l0,
                                                                  cutoff, tie, ambiguous_rel, (cplx*)Csite, (int*)bond,
                                                                  (int*)ambiguous, (cplx*)Tout, strides_from<7>(strides)));
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_chi2_bond(const void* L, int b, const void* T, const void* Bprev, int l0, double cutoff, double tie,
                            double ambiguous_rel, void* Csite, void* bond, void* ambiguous, void* Tout, void* stream) {
    return qm_chi2_bond_batch(L, b, T, Bprev, l0, cutoff, tie, ambiguous_rel, Csite, bond, ambiguous, Tout, 1, nullptr,
                              stream);
}

// sites: HOST array of n_sites device pointers to the (l, 2, r) tensors; dims: HOST int[n_sites + 1] bond sizes
// (dims[0] = dims[n_sites] = 1).  out: cplx[batch] = <0..0|psi> (not conjugated); tol < 0 skips the early-break check.
// strides: HOST long long[n_sites], batch stride of every site tensor (NULL for batch = 1); mismatch: int[batch].
extern "C" int qm_zero_overlap_batch(const void* const* sites, const int* dims, int n_sites, double tol, void* out,
                                     void* mismatch, int batch, const long long* strides, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_sites < 1 || n_sites > 32 || batch < 1) return -3;
    SiteList sl;
    for (int i = 0; i < n_sites; i++) {
        if (dims[i] > 1024 || dims[i + 1] > 1024) return -3;
        sl.p[i] = (const cplx*)sites[i];
        sl.l[i] = dims[i];
        sl.r[i] = dims[i + 1];
        sl.s[i] = strides ? strides[i] : 0;
    }
    QM_LAUNCH(QM_CLS_SMALL, st, k_zero_overlap<<<batch, NT, 0, st>>>(sl, n_sites, tol, (cplx*)out, (int*)mismatch));
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_zero_overlap(const void* const* sites, const int* dims, int n_sites, double tol, void* out,
                               void* mismatch, void* stream) {
    return qm_zero_overlap_batch(sites, dims, n_sites, tol, out, mismatch, 1, nullptr, stream);
}
