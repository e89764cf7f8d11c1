// Dense 2^N statevector kernels for the optimisation sweeps.
//
// Reference: qmprs/synthesis/mps_encoding/sequential.py
//   :215-292, :443-447  circuit tensor network contracted to a dense state   -> qm_circuit_state
//   :452-505            per gate: c <- G_old^H c ; E = tbar . c ; svd(E) ;
//                       G_new = conj(u vh) ; tbar <- tbar . G_new             -> qm_sweep
// Site i is axis i of the C-order reshape([2]*N): site i <-> bit (N-1-i) of the index.
// A two-site gate on (i, i+1) has matrix index 2*o_i + o_{i+1}; its low bit is q = N-2-i.
// These kernels are bandwidth-bound (HBM, or L2 when both vectors fit): every pass
// reads and writes each amplitude once with 16-byte accesses that are contiguous across
// a warp.
#include "common.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int NT = 256;

// op: 0 -> M = G, 1 -> M = G^H, 2 -> M = G^T
__device__ __forceinline__ void load_mat(const cplx* __restrict__ G, int dim, int op, cplx* M) {
    for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
            cplx v;
            if (op == 0) v = G[a * dim + b];
            else if (op == 1) v = cconj(G[b * dim + a]);
            else v = G[b * dim + a];
            M[a * dim + b] = v;
        }
}

// state <- M applied on the two bits (q+1, q).  One thread per group of 4 amplitudes.
__global__ void __launch_bounds__(NT)
k_gate2(cplx* __restrict__ x, int nbits, int q, const cplx* __restrict__ G, int op) {
    __shared__ cplx Ms[16];
    if (threadIdx.x == 0) load_mat(G, 4, op, Ms);
    __syncthreads();
    cplx M[16];
#pragma unroll
    for (int i = 0; i < 16; i++) M[i] = Ms[i];
    const long long ngroups = 1LL << (nbits - 2);
    const long long stride = 1LL << q;
    const long long lowmask = stride - 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ngroups;
         t += (long long)gridDim.x * blockDim.x) {
        long long base = ((t >> q) << (q + 2)) | (t & lowmask);
        cplx v0 = x[base], v1 = x[base + stride], v2 = x[base + 2 * stride], v3 = x[base + 3 * stride];
        cplx y[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            cplx s = cmul(M[a * 4 + 0], v0);
            cfma(s, M[a * 4 + 1], v1);
            cfma(s, M[a * 4 + 2], v2);
            cfma(s, M[a * 4 + 3], v3);
            y[a] = s;
        }
        x[base] = y[0]; x[base + stride] = y[1]; x[base + 2 * stride] = y[2]; x[base + 3 * stride] = y[3];
    }
}

// state <- M applied on bit q.
__global__ void __launch_bounds__(NT)
k_gate1(cplx* __restrict__ x, int nbits, int q, const cplx* __restrict__ G, int op) {
    __shared__ cplx Ms[4];
    if (threadIdx.x == 0) load_mat(G, 2, op, Ms);
    __syncthreads();
    cplx m0 = Ms[0], m1 = Ms[1], m2 = Ms[2], m3 = Ms[3];
    const long long ngroups = 1LL << (nbits - 1);
    const long long stride = 1LL << q;
    const long long lowmask = stride - 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ngroups;
         t += (long long)gridDim.x * blockDim.x) {
        long long base = ((t >> q) << (q + 1)) | (t & lowmask);
        cplx v0 = x[base], v1 = x[base + stride];
        x[base] = cadd(cmul(m0, v0), cmul(m1, v1));
        x[base + stride] = cadd(cmul(m2, v0), cmul(m3, v1));
    }
}

// ---------------------------------------------------------------------------------
// Polar factor of a d x d matrix (d = 2 or 4) by one-sided Jacobi:  E V = U Sigma,
// P = U V^H; null directions (sigma <= 1e-15 sigma_max, e.g. gates whose second input is
// still |0>) are completed to an orthonormal basis.  Writes conj(P) (sequential.py:478-491).
// Single thread.
// ---------------------------------------------------------------------------------
__device__ void polar_conj(const cplx* E, int d, cplx* out) {
    cplx A[4][4], V[4][4];
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) { A[i][j] = E[i * d + j]; V[i][j] = mk(i == j ? 1.0 : 0.0, 0.0); }
    const double tol2 = 4e-30;
    for (int sweep = 0; sweep < 40; sweep++) {
        int rot = 0;
        for (int p = 0; p < d - 1; p++)
            for (int q = p + 1; q < d; q++) {
                double a = 0.0, b = 0.0;
                cplx g = mk(0.0, 0.0);                     // g = a_p^H a_q
                for (int i = 0; i < d; i++) {
                    a += cabs2(A[i][p]); b += cabs2(A[i][q]);
                    ccfma(g, A[i][p], A[i][q]);
                }
                double mag2 = cabs2(g);
                if (!(a > 0.0 && b > 0.0) || mag2 <= tol2 * a * b) continue;
                rot = 1;
                double mag = sqrt(mag2);
                double zeta = (b - a) / (2.0 * mag);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                cplx e = mk(g.x / mag, g.y / mag);          // e^{i phi}
                cplx se = cscale(e, s), sec = cconj(se);
                // x' = c x - s e^{-i phi} y ; y' = s e^{i phi} x + c y
                for (int i = 0; i < d; i++) {
                    cplx xx = A[i][p], yy = A[i][q];
                    A[i][p] = csub(cscale(xx, c), cmul(sec, yy));
                    A[i][q] = cadd(cmul(se, xx), cscale(yy, c));
                    xx = V[i][p]; yy = V[i][q];
                    V[i][p] = csub(cscale(xx, c), cmul(sec, yy));
                    V[i][q] = cadd(cmul(se, xx), cscale(yy, c));
                }
            }
        if (!rot) break;
    }
    double sig[4], smax = 0.0;
    for (int j = 0; j < d; j++) {
        double s = 0.0;
        for (int i = 0; i < d; i++) s += cabs2(A[i][j]);
        sig[j] = sqrt(s);
        smax = sig[j] > smax ? sig[j] : smax;
    }
    bool isnull[4];
    for (int j = 0; j < d; j++) {
        isnull[j] = !(sig[j] > 1e-15 * smax) || smax == 0.0;
        if (!isnull[j]) {
            double inv = 1.0 / sig[j];
            for (int i = 0; i < d; i++) A[i][j] = cscale(A[i][j], inv);
        }
    }
    // complete null columns of U: Gram-Schmidt of the standard basis vector with the
    // largest residual against every column fixed so far (twice for orthogonality)
    bool fixed[4];
    for (int j = 0; j < d; j++) fixed[j] = !isnull[j];
    for (int j = 0; j < d; j++) {
        if (!isnull[j]) continue;
        double best = -1.0;
        cplx bestv[4];
        for (int k = 0; k < d; k++) {
            cplx v[4];
            for (int i = 0; i < d; i++) v[i] = mk(i == k ? 1.0 : 0.0, 0.0);
            for (int pass = 0; pass < 2; pass++)
                for (int c = 0; c < d; c++) {
                    if (!fixed[c]) continue;
                    cplx dot = mk(0.0, 0.0);
                    for (int i = 0; i < d; i++) ccfma(dot, A[i][c], v[i]);     // u_c^H v
                    for (int i = 0; i < d; i++) v[i] = csub(v[i], cmul(A[i][c], dot));
                }
            double nr = 0.0;
            for (int i = 0; i < d; i++) nr += cabs2(v[i]);
            if (nr > best * (1.0 + 1e-9)) {
                best = nr;
                for (int i = 0; i < d; i++) bestv[i] = v[i];
            }
        }
        double inv = 1.0 / sqrt(best);
        for (int i = 0; i < d; i++) A[i][j] = cscale(bestv[i], inv);
        fixed[j] = true;
    }
    // out = conj(U V^H):  out[i][j] = conj( sum_k U[i][k] conj(V[j][k]) )
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            cplx s = mk(0.0, 0.0);
            for (int k = 0; k < d; k++) cfmac(s, A[i][k], V[j][k]);
            out[i * d + j] = cconj(s);
        }
}

// ---------------------------------------------------------------------------------
// Environment tensor E[o][b] = sum_rest tbar[o,rest] * c[b,rest] over the gate's axes,
// followed (in the last CTA to finish) by the deterministic reduction of the per-CTA
// partials and the polar update of the gate in place.
// ---------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(NT)
k_env_polar(const cplx* __restrict__ tbar, const cplx* __restrict__ c, int nbits, int q,
            cplx* __restrict__ partials, unsigned int* __restrict__ counter, cplx* __restrict__ gate_out,
            cplx* __restrict__ env_out) {
    constexpr int K = (D == 4) ? 2 : 1;
    __shared__ double red[33];
    __shared__ int s_last;
    cplx acc[D * D];
#pragma unroll
    for (int i = 0; i < D * D; i++) acc[i] = mk(0.0, 0.0);
    const long long ngroups = 1LL << (nbits - K);
    const long long stride = 1LL << q;
    const long long lowmask = stride - 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ngroups;
         t += (long long)gridDim.x * blockDim.x) {
        long long base = ((t >> q) << (q + K)) | (t & lowmask);
        cplx tv[D], cv[D];
#pragma unroll
        for (int a = 0; a < D; a++) { tv[a] = tbar[base + a * stride]; cv[a] = c[base + a * stride]; }
#pragma unroll
        for (int o = 0; o < D; o++)
#pragma unroll
            for (int b = 0; b < D; b++) cfma(acc[o * D + b], tv[o], cv[b]);
    }
    // block reduction of 2*D*D doubles
    for (int i = 0; i < D * D; i++) {
        double re = block_sum(acc[i].x, red);
        double im = block_sum(acc[i].y, red);
        if (threadIdx.x == 0) partials[(long long)blockIdx.x * 16 + i] = mk(re, im);
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: fixed-order sum over CTAs, 256/(D*D) slices per entry then a serial combine
    __shared__ cplx Es[16];
    __shared__ cplx Ep[NT];
    {
        constexpr int E2 = D * D, NS = NT / E2;
        const int e = threadIdx.x % E2, sl = threadIdx.x / E2;
        cplx s = mk(0.0, 0.0);
        const volatile double* pv = (const volatile double*)partials;
        for (unsigned int b = sl; b < gridDim.x; b += NS) {
            long long o = ((long long)b * 16 + e) * 2;
            s.x += pv[o];
            s.y += pv[o + 1];
        }
        Ep[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x < E2) {
            cplx tsum = mk(0.0, 0.0);
            for (int k = 0; k < NS; k++) tsum = cadd(tsum, Ep[k * E2 + threadIdx.x]);
            Es[threadIdx.x] = tsum;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx E[16], P[16];
        for (int i = 0; i < D * D; i++) E[i] = Es[i];
        polar_conj(E, D, P);
        for (int i = 0; i < D * D; i++) gate_out[i] = P[i];
        if (env_out)
            for (int i = 0; i < D * D; i++) env_out[i] = E[i];
        *counter = 0u;
        __threadfence();
    }
}

__global__ void k_basis_state(cplx* __restrict__ x, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] = mk(i == 0 ? 1.0 : 0.0, 0.0);
}

int grid_groups(long long ngroups) {
    long long g = (ngroups + NT - 1) / NT;
    const long long cap = 148LL * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_gate(cplx* x, int nbits, int site, int kind, const cplx* G, int op, cudaStream_t st) {
    qm_prof_work(QM_CLS_GATE, 32.0 * (double)(1LL << nbits));      // read + write every amplitude
    if (kind == 2) {
        int q = nbits - 2 - site;
        QM_LAUNCH(QM_CLS_GATE, st, k_gate2<<<grid_groups(1LL << (nbits - 2)), NT, 0, st>>>(x, nbits, q, G, op));
    } else {
        int q = nbits - 1 - site;
        QM_LAUNCH(QM_CLS_GATE, st, k_gate1<<<grid_groups(1LL << (nbits - 1)), NT, 0, st>>>(x, nbits, q, G, op));
    }
    return (int)cudaGetLastError();
}

}  // namespace

// x <- op(G) x on site (kind 1) or sites (site, site+1) (kind 2).  G: device 2x2 / 4x4 row-major.
extern "C" int qm_apply_gate(void* x, int n_sites, int site, int kind, const void* G, int op, void* stream) {
    return launch_gate((cplx*)x, n_sites, site, kind, (const cplx*)G, op, (cudaStream_t)stream);
}

// c <- all gates applied in order to |0...0>.  gates: device [n_gates][16]; sites/kinds: host int arrays.
extern "C" int qm_circuit_state(void* c, int n_sites, const void* gates, const int* sites, const int* kinds,
                                int n_gates, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = 1LL << n_sites;
    QM_LAUNCH(QM_CLS_GATE, st, k_basis_state<<<grid_groups(n), NT, 0, st>>>((cplx*)c, n));
    for (int g = 0; g < n_gates; g++) {
        int e = launch_gate((cplx*)c, n_sites, sites[g], kinds[g], (const cplx*)gates + (long long)g * 16, 0, st);
        if (e) return e;
    }
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" long long qm_sweep_work_bytes(void) { return (long long)(148 * 8 * 16 * sizeof(cplx) + 256); }

// One environment sweep (sequential.py:452-505).  On entry c = circuit state with all
// gates applied, tbar = conj(target).  Gates are visited last-applied first and updated
// in place in `gates`.  envs (optional, device [n_gates][16]) receives each E.
extern "C" int qm_sweep(void* c_, void* tbar_, int n_sites, void* gates_, const int* sites, const int* kinds,
                        int n_gates, void* work, void* envs_, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cplx* c = (cplx*)c_;
    cplx* tbar = (cplx*)tbar_;
    cplx* gates = (cplx*)gates_;
    cplx* envs = (cplx*)envs_;
    unsigned int* counter = (unsigned int*)work;
    cplx* partials = (cplx*)((char*)work + 256);
    QM_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    for (int g = n_gates - 1; g >= 0; g--) {
        cplx* G = gates + (long long)g * 16;
        int e = launch_gate(c, n_sites, sites[g], kinds[g], G, 1, st);          // c <- G_old^H c
        if (e) return e;
        if (kinds[g] == 2) {
            int q = n_sites - 2 - sites[g];
            QM_LAUNCH(QM_CLS_ENV, st, (k_env_polar<4><<<grid_groups(1LL << (n_sites - 2)), NT, 0, st>>>(
                tbar, c, n_sites, q, partials, counter, G, envs ? envs + (long long)g * 16 : nullptr)));
        } else {
            int q = n_sites - 1 - sites[g];
            QM_LAUNCH(QM_CLS_ENV, st, (k_env_polar<2><<<grid_groups(1LL << (n_sites - 1)), NT, 0, st>>>(
                tbar, c, n_sites, q, partials, counter, G, envs ? envs + (long long)g * 16 : nullptr)));
        }
        qm_prof_work(QM_CLS_ENV, 32.0 * (double)(1LL << n_sites));               // read tbar and c
        e = launch_gate(tbar, n_sites, sites[g], kinds[g], G, 2, st);           // tbar <- G_new^T tbar
        if (e) return e;
    }
    QM_CHECK_LAUNCH();
    return 0;
}
