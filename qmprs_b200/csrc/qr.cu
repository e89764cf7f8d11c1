// Householder QR for complex128, row-major, LAPACK zgeqr2 / zlarfg conventions.
//
// Replaces quimb's qr_stabilized (numpy.linalg.qr + positive diagonal) reached from
// left_canonize / right_canonize / tensor_compress_bond, i.e. qmprs/primitives/mps.py:396-398
// and 451-453.
//
//   qm_qr       in place: upper triangle <- R, strictly-lower part <- reflector tails,
//               tau[j]; one (vector, apply) kernel pair per column, the trailing update
//               streams the remaining columns with coalesced 128-byte row segments.
//   qm_qr_formq explicit thin Q (m x k) = H_0 ... H_{k-1} [I;0]
//   qm_qr_posdiag  makes diag(R) non-negative (R row and Q column sign flips)
#include "common.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int NT = 256;
constexpr int CT = 8;    // columns per CTA in the trailing update
constexpr int RL = 32;   // row lanes per CTA

// Column j: x = A[j:, j].  zlarfg: beta = -sign(Re alpha)*||x||, tau = (beta-alpha)/beta,
// v = [1, x[1:]/(alpha-beta)].  Writes beta to A[j][j], v tail below it, tau[j].
__global__ void __launch_bounds__(NT)
k_house_vec(cplx* __restrict__ A, long long lda, int m, int j, cplx* __restrict__ tau) {
    __shared__ double red[33];
    __shared__ cplx s_scale;
    __shared__ int s_skip;
    const int tid = threadIdx.x;
    double part = 0.0;
    for (int r = j + 1 + tid; r < m; r += NT) part += cabs2(A[(long long)r * lda + j]);
    double xn2 = block_sum(part, red);
    if (tid == 0) {
        cplx alpha = A[(long long)j * lda + j];
        if (xn2 == 0.0 && alpha.y == 0.0) {
            tau[j] = mk(0.0, 0.0);
            s_skip = 1;
        } else {
            double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2);
            double beta = (alpha.x >= 0.0) ? -nrm : nrm;
            tau[j] = mk((beta - alpha.x) / beta, -alpha.y / beta);
            cplx d = mk(alpha.x - beta, alpha.y);           // alpha - beta
            double d2 = cabs2(d);
            s_scale = mk(d.x / d2, -d.y / d2);              // 1/(alpha-beta)
            A[(long long)j * lda + j] = mk(beta, 0.0);
            s_skip = 0;
        }
    }
    __syncthreads();
    if (s_skip) return;
    cplx sc = s_scale;
    for (int r = j + 1 + tid; r < m; r += NT) {
        cplx* p = A + (long long)r * lda + j;
        *p = cmul(*p, sc);
    }
}

// T[j:, c] -= coef * v * (v^H T[j:, c]) for columns c in [c_begin, ncols), v = [1, V[j+1:, j]],
// coef = conj(tau[j]) (factorisation, H^H from the left) or tau[j] (forming Q).
__global__ void __launch_bounds__(NT)
k_house_apply(cplx* __restrict__ T, long long ldt, int m, int ncols, int c_begin, const cplx* __restrict__ V,
              long long ldv, int j, const cplx* __restrict__ tau, int conj_tau) {
    __shared__ cplx part[RL][CT + 1];
    __shared__ cplx wv[CT];
    cplx tj = tau[j];
    if (tj.x == 0.0 && tj.y == 0.0) return;
    if (conj_tau) tj.y = -tj.y;
    const int tid = threadIdx.x;
    const int cl = tid % CT, rl = tid / CT;
    const int c = c_begin + blockIdx.x * CT + cl;
    const bool ok = c < ncols;
    cplx acc = mk(0.0, 0.0);
    if (ok) {
        for (int r = j + rl; r < m; r += RL) {
            cplx v = (r == j) ? mk(1.0, 0.0) : V[(long long)r * ldv + j];
            ccfma(acc, v, T[(long long)r * ldt + c]);
        }
    }
    part[rl][cl] = acc;
    __syncthreads();
    if (tid < CT) {
        cplx s = mk(0.0, 0.0);
        for (int i = 0; i < RL; i++) s = cadd(s, part[i][tid]);
        wv[tid] = cmul(tj, s);
    }
    __syncthreads();
    if (ok) {
        cplx w = wv[cl];
        for (int r = j + rl; r < m; r += RL) {
            cplx v = (r == j) ? mk(1.0, 0.0) : V[(long long)r * ldv + j];
            cplx* p = T + (long long)r * ldt + c;
            *p = csub(*p, cmul(v, w));
        }
    }
}

__global__ void k_set_eye(cplx* __restrict__ Q, long long ldq, int m, int k) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * k) return;
    int r = (int)(idx / k), c = (int)(idx % k);
    Q[(long long)r * ldq + c] = mk(r == c ? 1.0 : 0.0, 0.0);
}

// R (k x n upper trapezoid taken from the factored A) -> Rout with zeros below the
// diagonal and non-negative diagonal; Q columns get the matching sign.
__global__ void k_extract_r(const cplx* __restrict__ A, long long lda, int k, int n, cplx* __restrict__ R,
                            long long ldr) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)k * n) return;
    int r = (int)(idx / n), c = (int)(idx % n);
    cplx v = mk(0.0, 0.0);
    if (c >= r) {
        double d = A[(long long)r * lda + r].x;
        v = A[(long long)r * lda + c];
        if (d < 0.0) v = mk(-v.x, -v.y);
    }
    R[(long long)r * ldr + c] = v;
}

__global__ void k_q_signs(cplx* __restrict__ Q, long long ldq, int m, int k, const cplx* __restrict__ A,
                          long long lda) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * k) return;
    int r = (int)(idx / k), c = (int)(idx % k);
    if (A[(long long)c * lda + c].x < 0.0) {
        cplx* p = Q + (long long)r * ldq + c;
        *p = mk(-p->x, -p->y);
    }
}

// ---------------------------------------------------------------------------------
// Blocked Householder QR (compact WY, panels of NB columns): the panel is factored by the column kernels above,
// restricted to the panel; everything to its right is updated by three ZGEMMs on the FP64 tensor cores
// (k_zgemm_tma / DMMA):  C <- (I - V T V^H)^H C = C - V (T^H (V^H C)).   LAPACK zgeqrf / zlarft / zlarfb layout.
// ---------------------------------------------------------------------------------
constexpr int NB = 32;

// Vw (rows x jb, ld = jb) <- unit lower trapezoid of the panel starting at (j0, j0)
__global__ void k_extract_v(const cplx* __restrict__ A, long long lda, int rows, int j0, int jb, cplx* __restrict__ Vw) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * jb) return;
    const int r = (int)(idx / jb), c = (int)(idx % jb);
    cplx v = mk(0.0, 0.0);
    if (r == c) v = mk(1.0, 0.0);
    else if (r > c) v = A[(long long)(j0 + r) * lda + j0 + c];
    Vw[idx] = v;
}

// T (jb x jb, upper triangular) of H_0 ... H_{jb-1} = I - V T V^H from S = V^H V and tau (zlarft, forward,
// columnwise):  T[i][i] = tau_i,  T[0:i, i] = -tau_i T[0:i, 0:i] S[0:i, i].   One CTA.
__global__ void __launch_bounds__(NB * NB)
k_build_t(const cplx* __restrict__ S, const cplx* __restrict__ tau, int jb, cplx* __restrict__ T) {
    __shared__ cplx Ts[NB][NB + 1], Ss[NB][NB + 1];
    const int a = threadIdx.x / NB, b = threadIdx.x % NB;
    Ts[a][b] = mk(0.0, 0.0);
    Ss[a][b] = (a < jb && b < jb) ? S[a * jb + b] : mk(0.0, 0.0);
    __syncthreads();
    for (int i = 0; i < jb; i++) {
        const cplx ti = tau[i];
        if (b == 0 && a < i) {                             // thread (a, 0): entry T[a][i]
            cplx acc = mk(0.0, 0.0);
            for (int x = a; x < i; x++) cfma(acc, Ts[a][x], Ss[x][i]);
            Ts[a][i] = cmul(mk(-ti.x, -ti.y), acc);
        }
        if (threadIdx.x == 0) Ts[i][i] = ti;
        __syncthreads();
    }
    if (a < jb && b < jb) T[a * jb + b] = Ts[a][b];
}

struct QrWork { cplx* V; cplx* S; cplx* T; cplx* W1; cplx* W2; size_t total; };
QrWork qr_carve(int m, int n, void* base) {
    QrWork w;
    char* b = (char*)base;
    size_t off = 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.V = (cplx*)(b + off); off += al((size_t)m * NB * sizeof(cplx));
    w.S = (cplx*)(b + off); off += al((size_t)NB * NB * sizeof(cplx));
    w.T = (cplx*)(b + off); off += al((size_t)NB * NB * sizeof(cplx));
    w.W1 = (cplx*)(b + off); off += al((size_t)NB * n * sizeof(cplx));
    w.W2 = (cplx*)(b + off); off += al((size_t)NB * n * sizeof(cplx));
    w.total = off;
    return w;
}

// V, T of the panel [j0, j0 + jb) of the factored A
int panel_vt(const cplx* A, long long lda, int m, int j0, int jb, const cplx* tau, const QrWork& w, cudaStream_t st) {
    const int rows = m - j0;
    QM_LAUNCH(QM_CLS_QR_VEC, st, k_extract_v<<<ceil_div((long long)rows * jb, 256), 256, 0, st>>>(A, lda, rows, j0, jb, w.V));
    int e = qm_zgemm(jb, jb, rows, 1.0, 0.0, w.V, jb, w.V, jb, 0.0, 0.0, w.S, jb, 1, 0, 0, 0, 1, (void*)st);   // S = V^H V
    if (e) return e;
    QM_LAUNCH(QM_CLS_QR_VEC, st, k_build_t<<<1, NB * NB, 0, st>>>(w.S, tau + j0, jb, w.T));
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" long long qm_qr_work_bytes(int m, int n) {
    return (long long)qr_carve(m, n > m ? n : m, nullptr).total;
}

// Blocked in-place Householder QR of A (m x n, lda): same output layout as qm_qr (R above the diagonal, reflector
// tails below, tau), trailing updates on the FP64 tensor cores.  work: qm_qr_work_bytes(m, n) bytes.
extern "C" int qm_qr_blocked(int m, int n, void* A_, long long lda, void* tau_, void* work, long long work_bytes,
                             void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    cplx* A = (cplx*)A_;
    cplx* tau = (cplx*)tau_;
    const int k = m < n ? m : n;
    QrWork w = qr_carve(m, n > m ? n : m, work);
    if ((long long)w.total > work_bytes) return -1;
    for (int j0 = 0; j0 < k; j0 += NB) {
        const int jb = (j0 + NB < k) ? NB : k - j0;
        // panel: column kernels restricted to columns [j0, j0 + jb)
        for (int j = j0; j < j0 + jb; j++) {
            QM_LAUNCH(QM_CLS_QR_VEC, st, k_house_vec<<<1, NT, 0, st>>>(A, lda, m, j, tau));
            const int ntrail = j0 + jb - (j + 1);
            if (ntrail > 0)
                QM_LAUNCH(QM_CLS_QR_APPLY, st, k_house_apply<<<ceil_div(ntrail, CT), NT, 0, st>>>(A, lda, m, j0 + jb, j + 1, A, lda, j, tau, 1));
        }
        const int nt = n - (j0 + jb);
        if (nt <= 0) continue;
        int e = panel_vt(A, lda, m, j0, jb, tau, w, st);
        if (e) return e;
        const int rows = m - j0;
        cplx* C = A + (long long)j0 * lda + j0 + jb;
        // W1 = V^H C ; W2 = T^H W1 ; C -= V W2
        e = qm_zgemm(jb, nt, rows, 1.0, 0.0, w.V, jb, C, lda, 0.0, 0.0, w.W1, nt, 1, 0, 0, 0, 1, stream_);
        if (e) return e;
        e = qm_zgemm(jb, nt, jb, 1.0, 0.0, w.T, jb, w.W1, nt, 0.0, 0.0, w.W2, nt, 1, 0, 0, 0, 1, stream_);
        if (e) return e;
        e = qm_zgemm(rows, nt, jb, -1.0, 0.0, w.V, jb, w.W2, nt, 1.0, 0.0, C, lda, 1, 0, 0, 0, 0, stream_);
        if (e) return e;
        qm_prof_work(QM_CLS_QR_APPLY, 16.0 * rows * (double)nt * jb);
    }
    QM_CHECK_LAUNCH();
    return 0;
}

// Q (m x k, ldq) = H_0 ... H_{k-1} [I; 0] from the factored A and tau, panel by panel from the last:
// Q[j0:, j0:] <- (I - V T V^H) Q[j0:, j0:].
extern "C" int qm_qr_formq_blocked(int m, int k, const void* A_, long long lda, const void* tau_, void* Q_, long long ldq,
                                   void* work, long long work_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    const cplx* A = (const cplx*)A_;
    const cplx* tau = (const cplx*)tau_;
    cplx* Q = (cplx*)Q_;
    QrWork w = qr_carve(m, k > m ? k : m, work);
    if ((long long)w.total > work_bytes) return -1;
    QM_LAUNCH(QM_CLS_SMALL, st, k_set_eye<<<ceil_div((long long)m * k, 256), 256, 0, st>>>(Q, ldq, m, k));
    const int npanels = (k + NB - 1) / NB;
    for (int p = npanels - 1; p >= 0; p--) {
        const int j0 = p * NB, jb = (j0 + NB < k) ? NB : k - j0;
        int e = panel_vt(A, lda, m, j0, jb, tau, w, st);
        if (e) return e;
        const int rows = m - j0, nc = k - j0;
        cplx* C = Q + (long long)j0 * ldq + j0;
        // W1 = V^H C ; W2 = T W1 ; C -= V W2
        e = qm_zgemm(jb, nc, rows, 1.0, 0.0, w.V, jb, C, ldq, 0.0, 0.0, w.W1, nc, 1, 0, 0, 0, 1, stream_);
        if (e) return e;
        e = qm_zgemm(jb, nc, jb, 1.0, 0.0, w.T, jb, w.W1, nc, 0.0, 0.0, w.W2, nc, 1, 0, 0, 0, 0, stream_);
        if (e) return e;
        e = qm_zgemm(rows, nc, jb, -1.0, 0.0, w.V, jb, w.W2, nc, 1.0, 0.0, C, ldq, 1, 0, 0, 0, 0, stream_);
        if (e) return e;
        qm_prof_work(QM_CLS_QR_APPLY, 16.0 * rows * (double)nc * jb);
    }
    QM_CHECK_LAUNCH();
    return 0;
}

// In-place Householder QR of A (m x n, lda).  tau: min(m,n) complex.
extern "C" int qm_qr(int m, int n, void* A_, long long lda, void* tau_, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    cplx* A = (cplx*)A_;
    cplx* tau = (cplx*)tau_;
    int k = m < n ? m : n;
    for (int j = 0; j < k; j++) {
        QM_LAUNCH(QM_CLS_QR_VEC, st, k_house_vec<<<1, NT, 0, st>>>(A, lda, m, j, tau));
        int ntrail = n - (j + 1);
        qm_prof_work(QM_CLS_QR_APPLY, 16.0 * (double)(m - j) * (ntrail > 0 ? ntrail : 0));   // dot + axpy
        if (ntrail > 0)
            QM_LAUNCH(QM_CLS_QR_APPLY, st, k_house_apply<<<ceil_div(ntrail, CT), NT, 0, st>>>(A, lda, m, n, j + 1, A, lda, j, tau, 1));
    }
    QM_CHECK_LAUNCH();
    return 0;
}

// Q (m x k, ldq) from the factored A and tau.
extern "C" int qm_qr_formq(int m, int k, const void* A_, long long lda, const void* tau_, void* Q_, long long ldq,
                           void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    cplx* Q = (cplx*)Q_;
    QM_LAUNCH(QM_CLS_SMALL, st, k_set_eye<<<ceil_div((long long)m * k, 256), 256, 0, st>>>(Q, ldq, m, k));
    for (int j = k - 1; j >= 0; j--) {
        // H_j only touches rows >= j; columns < j of the accumulated product are still e_c there (zero)
        int c0 = j;
        QM_LAUNCH(QM_CLS_QR_APPLY, st, k_house_apply<<<ceil_div(k - c0, CT), NT, 0, st>>>(Q, ldq, m, k, c0, (const cplx*)A_, lda, j,
                                                            (const cplx*)tau_, 0));
    }
    QM_CHECK_LAUNCH();
    return 0;
}

// R (k x n, ldr) with non-negative diagonal; flips the matching columns of Q (may be NULL).
extern "C" int qm_qr_finish(int m, int n, const void* A_, long long lda, void* R_, long long ldr, void* Q_,
                            long long ldq, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    int k = m < n ? m : n;
    QM_LAUNCH(QM_CLS_SMALL, st, k_extract_r<<<ceil_div((long long)k * n, 256), 256, 0, st>>>((const cplx*)A_, lda, k, n, (cplx*)R_, ldr));
    if (Q_)
        QM_LAUNCH(QM_CLS_SMALL, st, k_q_signs<<<ceil_div((long long)m * k, 256), 256, 0, st>>>((cplx*)Q_, ldq, m, k, (const cplx*)A_, lda));
    QM_CHECK_LAUNCH();
    return 0;
}
