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
#include "small_linalg.cuh"
#include "polar.cuh"
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
k_gate2(const cplx* xin, cplx* x, int nbits, int q, const cplx* __restrict__ G, int op) {
    pdl_wait();
    pdl_trigger();
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
        cplx v0 = xin[base], v1 = xin[base + stride], v2 = xin[base + 2 * stride], v3 = xin[base + 3 * stride];
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
k_gate1(const cplx* xin, cplx* x, int nbits, int q, const cplx* __restrict__ G, int op) {
    pdl_wait();
    pdl_trigger();
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
        cplx v0 = xin[base], v1 = xin[base + stride];
        x[base] = cadd(cmul(m0, v0), cmul(m1, v1));
        x[base + stride] = cadd(cmul(m2, v0), cmul(m3, v1));
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


// ---------------------------------------------------------------------------------
// Stored-intermediates sweep.  While the circuit state is built (A8) every intermediate
// c_k = g_{k-1}...g_0|0> is kept in HBM ((M+1) x 2^N amplitudes: 4.8 GB at the 20-qubit
// headline config, 180 GB available), so the backward sweep never re-derives c:
// one fused pass per gate reads tbar, applies the previously updated gate (its axes sit
// directly below the current gate's axes), writes tbar, reads c_k and accumulates E_k.
// 48 bytes per amplitude per gate-step instead of 96.
// ---------------------------------------------------------------------------------
// NU: bits in the window [q0, q0+NU); the current gate (dimension CD) acts on the top
// log2(CD) bits of the window, the pending gate (dimension PD, 0 = none) on the bottom ones.
// Common tail of the environment kernels: warp transpose-reduce of the 32 accumulators, CTA partial,
// ticket; the last CTA sums the partials in fixed order and does the polar update of the gate.
template <int CD>
__device__ __forceinline__ void env_epilogue(double* acc, cplx* __restrict__ partials,
                                             unsigned int* __restrict__ counter, cplx* __restrict__ gate_out,
                                             cplx* __restrict__ env_out, cplx* __restrict__ vwarm) {
    __shared__ double wsum[NT / 32][32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double mine = warp_reduce32(acc, lane);
    wsum[warp][lane] = mine;
    __syncthreads();
    if (warp == 0) {
        double ssum = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) ssum += wsum[w][lane];
        ((double*)partials)[(long long)blockIdx.x * 32 + lane] = ssum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                     // cumulative: publishes the CTA's partials before the ticket
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: fixed-order reduction over CTAs (8 slices per value, then serial combine), then the polar update
    __shared__ double Ep[NT];
    __shared__ cplx Es[16];
    {
        const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
        // L2 loads (other CTAs wrote these; the fence above orders them), batched so they overlap
        const double* pv = (const double*)partials;
        double ssum = 0.0;
#pragma unroll 8
        for (unsigned int b = sl; b < gridDim.x; b += NT / 32) ssum += __ldcg(pv + (long long)b * 32 + e);
        Ep[threadIdx.x] = ssum;
        __syncthreads();
        if (threadIdx.x < 32) {
            double tsum = 0.0;
#pragma unroll
            for (int k = 0; k < NT / 32; k++) tsum += Ep[k * 32 + threadIdx.x];
            Ep[threadIdx.x] = tsum;
        }
        __syncthreads();
        if (threadIdx.x < CD * CD) Es[threadIdx.x] = mk(Ep[2 * threadIdx.x], Ep[2 * threadIdx.x + 1]);
    }
    __syncthreads();
    __shared__ cplx pol_scratch[32];
    if (threadIdx.x < 32) {
        polar_conj_warp(Es, CD, gate_out, pol_scratch, vwarm, vwarm);
        if (env_out && threadIdx.x < CD * CD) env_out[threadIdx.x] = Es[threadIdx.x];
        if (threadIdx.x == 0) {
            *counter = 0u;
            __threadfence();
        }
    }
}

template <int NU, int CD, int PD>
__global__ void __launch_bounds__(NT)
k_env_fused(cplx* __restrict__ tbar, const cplx* __restrict__ c, int nbits, int q0,
            const cplx* __restrict__ Gpend, cplx* __restrict__ partials, unsigned int* __restrict__ counter,
            cplx* __restrict__ gate_out, cplx* __restrict__ env_out, cplx* __restrict__ vwarm) {
    constexpr int GSZ = 1 << NU;
    constexpr int CSH = NU - (CD == 4 ? 2 : 1);
    constexpr int NLOW = 1 << CSH;
    constexpr int PDD = PD > 0 ? PD : 1;     // divisor that is never zero in the PD == 0 instantiations
    __shared__ cplx Pm[16];
    const long long ngroups = 1LL << (nbits - NU);
    const long long stride = 1LL << q0;
    const long long lowmask = stride - 1;
    // The stored circuit state c_k was written by the forward pass, long before the predecessor of this
    // kernel started (the predecessor triggers only after its own wait): its loads can be in flight while
    // the predecessor finishes its reduction and polar update.  tbar and the pending gate cannot.
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    cplx cv[GSZ];
    if (t < ngroups) {
        const long long base = ((t >> q0) << (q0 + NU)) | (t & lowmask);
#pragma unroll
        for (int j = 0; j < GSZ; j++) cv[j] = c[base + j * stride];
    }
    pdl_wait();
    pdl_trigger();
    if (PD > 0 && threadIdx.x < PD * PD) {
        int a = threadIdx.x / PDD, b = threadIdx.x % PDD;
        Pm[threadIdx.x] = Gpend[b * PD + a];                  // M = G^T : tbar'[b] = sum_o G[o][b] tbar[o]
    }
    __syncthreads();
    const cplx* P = Pm;                      // broadcast reads from shared memory (keeps 64 registers free)
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    bool first = true;
    for (; t < ngroups; t += (long long)gridDim.x * blockDim.x) {
        const long long base = ((t >> q0) << (q0 + NU)) | (t & lowmask);
        cplx tv[GSZ];
#pragma unroll
        for (int j = 0; j < GSZ; j++) tv[j] = tbar[base + j * stride];
        if (!first) {
#pragma unroll
            for (int j = 0; j < GSZ; j++) cv[j] = c[base + j * stride];
        }
        first = false;
        if (PD > 0) {
#pragma unroll
            for (int h = 0; h < GSZ / PDD; h++) {
                cplx y[PD > 0 ? PD : 1];
#pragma unroll
                for (int a = 0; a < PD; a++) {
                    cplx sacc = mk(0.0, 0.0);
#pragma unroll
                    for (int b = 0; b < PD; b++) cfma(sacc, P[a * PD + b], tv[h * PD + b]);
                    y[a] = sacc;
                }
#pragma unroll
                for (int a = 0; a < PD; a++) tv[h * PD + a] = y[a];
            }
#pragma unroll
            for (int j = 0; j < GSZ; j++) tbar[base + j * stride] = tv[j];
        }
#pragma unroll
        for (int o = 0; o < CD; o++)
#pragma unroll
            for (int b = 0; b < CD; b++) {
                cplx e = mk(acc[2 * (o * CD + b)], acc[2 * (o * CD + b) + 1]);
#pragma unroll
                for (int lo = 0; lo < NLOW; lo++) cfma(e, tv[(o << CSH) | lo], cv[(b << CSH) | lo]);
                acc[2 * (o * CD + b)] = e.x;
                acc[2 * (o * CD + b) + 1] = e.y;
            }
    }
    env_epilogue<CD>(acc, partials, counter, gate_out, env_out, vwarm);
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

int launch_gate(cplx* x, int nbits, int site, int kind, const cplx* G, int op, cudaStream_t st,
                const cplx* xin = nullptr) {
    if (!xin) xin = x;
    qm_prof_work(QM_CLS_GATE, 32.0 * (double)(1LL << nbits));      // read + write every amplitude
    if (kind == 2) {
        int q = nbits - 2 - site;
        QM_LAUNCH(QM_CLS_GATE, st, qm_launch_dep(k_gate2, dim3(grid_groups(1LL << (nbits - 2))), dim3(NT), 0, st,
                                                 xin, x, nbits, q, G, op));
    } else {
        int q = nbits - 1 - site;
        QM_LAUNCH(QM_CLS_GATE, st, qm_launch_dep(k_gate1, dim3(grid_groups(1LL << (nbits - 1))), dim3(NT), 0, st,
                                                 xin, x, nbits, q, G, op));
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


// cs: (n_gates+1) x 2^N amplitudes; cs[0] = |0..0>, cs[k+1] = g_k cs[k].  A8 with every
// intermediate kept for the stored sweep.
extern "C" int qm_circuit_states(void* cs_, int n_sites, const void* gates, const int* sites, const int* kinds,
                                 int n_gates, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cplx* cs = (cplx*)cs_;
    const long long n = 1LL << n_sites;
    QM_LAUNCH(QM_CLS_GATE, st, k_basis_state<<<grid_groups(n), NT, 0, st>>>(cs, n));
    for (int g = 0; g < n_gates; g++) {
        int e = launch_gate(cs + (long long)(g + 1) * n, n_sites, sites[g], kinds[g],
                            (const cplx*)gates + (long long)g * 16, 0, st, cs + (long long)g * n);
        if (e) return e;
    }
    QM_CHECK_LAUNCH();
    return 0;
}

namespace {
template <int NU, int CD, int PD>
void launch_env_fused(cplx* tbar, const cplx* c, int nbits, int q0, const cplx* Gpend, cplx* partials,
                      unsigned int* counter, cplx* gate_out, cplx* env_out, cplx* vwarm, cudaStream_t st) {
    const long long ngroups = 1LL << (nbits - NU);
    qm_prof_work(QM_CLS_ENV, (PD > 0 ? 48.0 : 32.0) * (double)(1LL << nbits));
    // one group per thread (a 148-CTA persistent grid measured slower: 35 vs 30 us at 20 qubits, the
    // per-iteration load latency is not overlapped at 1 CTA/SM)
    QM_LAUNCH(QM_CLS_ENV, st, qm_launch_dep(k_env_fused<NU, CD, PD>, dim3(grid_groups(ngroups)), dim3(NT), 0, st,
                                            tbar, c, nbits, q0, Gpend, partials, counter, gate_out, env_out, vwarm));
}
}  // namespace

// One environment sweep using the stored intermediates cs (from qm_circuit_states).
// tbar = conj(target) on entry; gates updated in place.
extern "C" int qm_sweep_stored(const void* cs_, void* tbar_, int n_sites, void* gates_, const int* sites,
                               const int* kinds, int n_gates, void* work, void* envs_, void* vwarm_, void* stream) {
    cplx* vwarm = (cplx*)vwarm_;
    cudaStream_t st = (cudaStream_t)stream;
    const cplx* cs = (const cplx*)cs_;
    cplx* tbar = (cplx*)tbar_;
    cplx* gates = (cplx*)gates_;
    cplx* envs = (cplx*)envs_;
    unsigned int* counter = (unsigned int*)work;
    cplx* partials = (cplx*)((char*)work + 256);
    const long long n = 1LL << n_sites;
    const int N = n_sites;
    QM_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    for (int g = n_gates - 1; g >= 0; g--) {
        cplx* G = gates + (long long)g * 16;
        cplx* env = envs ? envs + (long long)g * 16 : nullptr;
        const cplx* c = cs + (long long)g * n;
        cplx* vw = vwarm ? vwarm + (long long)g * 16 : nullptr;
        const int ck = kinds[g];
        const int cb = (ck == 2) ? N - 2 - sites[g] : N - 1 - sites[g];      // lowest bit of the current gate
        bool fused = false;
        if (g + 1 < n_gates) {
            const cplx* Gp = gates + (long long)(g + 1) * 16;
            const int pk = kinds[g + 1];
            const int pb = (pk == 2) ? N - 2 - sites[g + 1] : N - 1 - sites[g + 1];
            const int ptop = pb + (pk == 2 ? 1 : 0);                          // highest bit of the pending gate
            if (ck == 2 && pk == 2 && ptop == cb) {                          // overlap on one site
                // (a two-threads-per-group variant with half the registers measured slower: 50 vs 30 us at 20 qubits)
                launch_env_fused<3, 4, 4>(tbar, c, N, pb, Gp, partials, counter, G, env, vw, st);
                fused = true;
            } else if (ck == 2 && pk == 1 && pb == cb) {
                launch_env_fused<2, 4, 2>(tbar, c, N, pb, Gp, partials, counter, G, env, vw, st);
                fused = true;
            } else if (ck == 1 && pk == 2 && ptop == cb - 1) {
                launch_env_fused<3, 2, 4>(tbar, c, N, pb, Gp, partials, counter, G, env, vw, st);
                fused = true;
            } else if (ck == 1 && pk == 1 && pb == cb - 1) {
                launch_env_fused<2, 2, 2>(tbar, c, N, pb, Gp, partials, counter, G, env, vw, st);
                fused = true;
            } else {
                int e = launch_gate(tbar, N, sites[g + 1], pk, Gp, 2, st);    // tbar <- G_new^T tbar (unfused)
                if (e) return e;
            }
        }
        if (!fused) {
            if (ck == 2) launch_env_fused<2, 4, 0>(tbar, c, N, cb, nullptr, partials, counter, G, env, vw, st);
            else launch_env_fused<1, 2, 0>(tbar, c, N, cb, nullptr, partials, counter, G, env, vw, st);
        }
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
