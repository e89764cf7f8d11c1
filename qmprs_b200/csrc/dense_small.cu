// Small registers: all optimisation sweeps of a state in one launch (SURVEY 8a rows A8 + A9 for states whose two
// dense vectors fit one SM's shared memory).  Reference: qmprs/synthesis/mps_encoding/sequential.py:443-505, 509-541.
// Own translation unit: see polar.cuh.
#include "common.cuh"
#include "polar.cuh"
#include "qmprs_b200.h"

namespace {

// ---------------------------------------------------------------------------------
// Small registers (2^N amplitudes x 2 vectors fit one SM's shared memory: N <= 12): ALL optimisation
// sweeps of a state in ONE launch, one CTA per state, grid = batch ("12 q batch: per state
// 2 400 gate-steps x 256 KB, SMEM-resident").  Same arithmetic as qm_circuit_state + qm_sweep
// (sequential.py:443-505): per sweep  c = all gates on |0..0>,  tbar = conj(target);  per gate, last first:
//   c <- G_old^H c ;  E = sum_rest tbar . c ;  G_new = conj(polar(E)) ;  tbar <- G_new^T tbar.
// The c update and the environment accumulation share one pass; the 32 partial sums are reduced with a warp
// transpose-reduce and a fixed-order sum over warps (deterministic); the 4x4 polar runs in warp 0.
// Replaces ~250 launches per sweep (5000 of the 11 500 graph nodes of a 12-qubit / 10-layer / 20-sweep state).
// ---------------------------------------------------------------------------------
constexpr int NTS = 512;

template <int D, int OP>      // OP as load_mat: 0 -> G, 1 -> G^H, 2 -> G^T;  matrix read from shared memory (broadcast)
__device__ __forceinline__ cplx gate_elem(const cplx* G, int a, int b) {
    if (OP == 0) return G[a * D + b];
    if (OP == 1) return cconj(G[b * D + a]);
    return G[b * D + a];
}

template <int D, int OP>
__device__ __forceinline__ void small_apply(cplx* x, int nst, int q, const cplx* G) {
    constexpr int K = (D == 4) ? 2 : 1;
    const int ngroups = nst >> K, stride = 1 << q, lowmask = stride - 1;
    for (int tg = threadIdx.x; tg < ngroups; tg += blockDim.x) {
        const int base = ((tg >> q) << (q + K)) | (tg & lowmask);
        cplx v[D], y[D];
#pragma unroll
        for (int a = 0; a < D; a++) v[a] = x[base + a * stride];
#pragma unroll
        for (int a = 0; a < D; a++) {
            cplx sacc = cmul(gate_elem<D, OP>(G, a, 0), v[0]);
#pragma unroll
            for (int b = 1; b < D; b++) cfma(sacc, gate_elem<D, OP>(G, a, b), v[b]);
            y[a] = sacc;
        }
#pragma unroll
        for (int a = 0; a < D; a++) x[base + a * stride] = y[a];
    }
}

// c <- G^H c on the gate's bits, acc[2(oD+b)] += tbar[o,rest] c_new[b,rest]
template <int D>
__device__ __forceinline__ void small_env(cplx* c, const cplx* t, int nst, int q, const cplx* G, double* acc) {
    constexpr int K = (D == 4) ? 2 : 1;
    const int ngroups = nst >> K, stride = 1 << q, lowmask = stride - 1;
    for (int tg = threadIdx.x; tg < ngroups; tg += blockDim.x) {
        const int base = ((tg >> q) << (q + K)) | (tg & lowmask);
        cplx v[D], y[D], tv[D];
#pragma unroll
        for (int a = 0; a < D; a++) { v[a] = c[base + a * stride]; tv[a] = t[base + a * stride]; }
#pragma unroll
        for (int a = 0; a < D; a++) {
            cplx sacc = cmul(gate_elem<D, 1>(G, a, 0), v[0]);
#pragma unroll
            for (int b = 1; b < D; b++) cfma(sacc, gate_elem<D, 1>(G, a, b), v[b]);
            y[a] = sacc;
        }
#pragma unroll
        for (int a = 0; a < D; a++) c[base + a * stride] = y[a];
#pragma unroll
        for (int o = 0; o < D; o++)
#pragma unroll
            for (int b = 0; b < D; b++) {
                cplx e = mk(acc[2 * (o * D + b)], acc[2 * (o * D + b) + 1]);
                cfma(e, tv[o], y[b]);
                acc[2 * (o * D + b)] = e.x;
                acc[2 * (o * D + b) + 1] = e.y;
            }
    }
}

__global__ void __launch_bounds__(NTS, 1)
k_sweeps_small(const cplx* __restrict__ targets, int nbits, cplx* __restrict__ gates_g, const int* __restrict__ sites,
               const int* __restrict__ kinds, int n_gates, int num_sweeps, cplx* __restrict__ envs_g, int warm,
               const cplx* __restrict__ psis, double* __restrict__ overlaps) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    const int nst = 1 << nbits;
    cplx* c = (cplx*)sw_smem;
    cplx* t = c + nst;
    cplx* g = t + nst;                                   // [n_gates][16]
    // warm start of each gate's 4x4 Jacobi polar: right singular vectors of the previous sweep (as qm_sweep_stored)
    cplx* vw = warm ? g + (long long)n_gates * 16 : nullptr;
    int* gq = (int*)(g + (long long)n_gates * 16 * (warm ? 2 : 1));     // lowest bit of each gate
    int* gd = gq + n_gates;                              // dimension (2 or 4)
    __shared__ double wsum[NTS / 32][32];
    __shared__ cplx Es[16];
    __shared__ cplx pol_scratch[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const cplx* target = targets + (long long)blockIdx.x * nst;
    cplx* gg = gates_g + (long long)blockIdx.x * n_gates * 16;
    cplx* envs = envs_g ? envs_g + (long long)blockIdx.x * n_gates * 16 : nullptr;
    for (int i = tid; i < n_gates * 16; i += blockDim.x) {
        g[i] = gg[i];
        if (warm) vw[i] = mk(0.0, 0.0);
    }
    for (int k = tid; k < n_gates; k += blockDim.x) {
        const int kd = kinds[k];
        gd[k] = kd == 2 ? 4 : 2;
        gq[k] = kd == 2 ? nbits - 2 - sites[k] : nbits - 1 - sites[k];
    }
    __syncthreads();
    for (int sweep = 0; sweep < num_sweeps; sweep++) {
        for (int i = tid; i < nst; i += blockDim.x) c[i] = mk(i == 0 ? 1.0 : 0.0, 0.0);
        __syncthreads();
        for (int k = 0; k < n_gates; k++) {
            if (gd[k] == 4) small_apply<4, 0>(c, nst, gq[k], g + k * 16);
            else small_apply<2, 0>(c, nst, gq[k], g + k * 16);
            __syncthreads();
        }
        for (int i = tid; i < nst; i += blockDim.x) t[i] = cconj(target[i]);
        __syncthreads();
        for (int k = n_gates - 1; k >= 0; k--) {
            const int d = gd[k], q = gq[k];
            cplx* G = g + k * 16;
            double acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.0;
            if (d == 4) small_env<4>(c, t, nst, q, G, acc);
            else small_env<2>(c, t, nst, q, G, acc);
            const double mine = warp_reduce32(acc, lane);
            wsum[warp][lane] = mine;
            __syncthreads();
            if (warp == 0) {
                double ssum = 0.0;
                for (int w = 0; w < nw; w++) ssum += wsum[w][lane];     // fixed order
                // lane holds one double of E: (re, im) of entry lane/2
                const double other = __shfl_xor_sync(0xffffffffu, ssum, 1);
                if ((lane & 1) == 0 && (lane >> 1) < d * d) Es[lane >> 1] = mk(ssum, other);
                __syncwarp();
                polar_conj_warp(Es, d, G, pol_scratch, warm ? vw + k * 16 : nullptr, warm ? vw + k * 16 : nullptr);
                // the rank-deficient branch of the polar returns early in 31 lanes while lane 0 finishes the
                // single-thread completion: reconverge before the block barrier
                __syncwarp();
                if (envs && sweep == num_sweeps - 1 && lane < d * d) envs[k * 16 + lane] = Es[lane];
            }
            __syncthreads();
            if (d == 4) small_apply<4, 2>(t, nst, q, G);
            else small_apply<2, 2>(t, nst, q, G);
            __syncthreads();
        }
    }
    for (int i = tid; i < n_gates * 16; i += blockDim.x) gg[i] = g[i];
    if (overlaps) {
        // <psi | circuit |0..0>> / |psi| with the final gates (psi: the state's own row of `psis`, else the target):
        // the fidelity users report (README.md:66), so that a batch needs no per-state circuit + dot kernels
        const cplx* psi = psis ? psis + (long long)blockIdx.x * nst : target;
        __syncthreads();
        for (int i = tid; i < nst; i += blockDim.x) c[i] = mk(i == 0 ? 1.0 : 0.0, 0.0);
        __syncthreads();
        for (int k = 0; k < n_gates; k++) {
            if (gd[k] == 4) small_apply<4, 0>(c, nst, gq[k], g + k * 16);
            else small_apply<2, 0>(c, nst, gq[k], g + k * 16);
            __syncthreads();
        }
        cplx ov = mk(0.0, 0.0);
        double nr = 0.0;
        for (int i = tid; i < nst; i += blockDim.x) {
            const cplx p = psi[i];
            ccfma(ov, p, c[i]);
            nr += cabs2(p);
        }
        double v3[3] = {ov.x, ov.y, nr};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            v3[j] = warp_sum(v3[j]);
            if (lane == 0) wsum[warp][j] = v3[j];
        }
        __syncthreads();
        if (tid == 0) {
            double s3[3] = {0.0, 0.0, 0.0};
            for (int w = 0; w < nw; w++)                              // fixed order
                for (int j = 0; j < 3; j++) s3[j] += wsum[w][j];
            const double inv = s3[2] > 0.0 ? rsqrt(s3[2]) : 0.0;
            overlaps[2 * blockIdx.x] = s3[0] * inv;
            overlaps[2 * blockIdx.x + 1] = s3[1] * inv;
        }
    }
}

}  // namespace

// All `num_sweeps` optimisation sweeps of `batch` independent small states (one CTA each; n_sites <= 12 and
// n_gates <= 256 so that 2 vectors + the gates fit in shared memory).  targets: device [batch][2^N] (NOT conjugated),
// gates: device [batch][n_gates][16] in application order, updated in place; sites/kinds: DEVICE int[n_gates]
// (one schedule for the whole batch); envs (optional): [batch][n_gates][16] environments of the last sweep.
// Returns -3 when the state does not fit (callers then use qm_circuit_states + qm_sweep_stored).
// psis / overlaps (optional, both may be NULL): overlaps[batch][2] <- <psi_b| circuit_b |0..0> / |psi_b| with the final
// gates (psi_b = psis[b] or the target when psis is NULL); with overlaps given, num_sweeps = 0 is allowed.
extern "C" int qm_sweeps_small(const void* targets, int n_sites, void* gates, const int* sites_dev, const int* kinds_dev,
                               int n_gates, int num_sweeps, int batch, void* envs, const void* psis, void* overlaps,
                               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_sites < 2 || n_sites > 12 || n_gates < 1 || n_gates > 256 || batch < 1) return -3;
    if (num_sweeps <= 0 && !overlaps) return 0;
    if (num_sweeps < 0) num_sweeps = 0;
    const size_t nst = (size_t)1 << n_sites;
    size_t smem = 2 * nst * sizeof(cplx) + (size_t)n_gates * 16 * sizeof(cplx) + (size_t)n_gates * 2 * sizeof(int);
    const int warm = smem + (size_t)n_gates * 16 * sizeof(cplx) <= 220 * 1024;      // room for the polar warm starts
    if (warm) smem += (size_t)n_gates * 16 * sizeof(cplx);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        QM_CUDA(cudaFuncSetAttribute(k_sweeps_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    long long groups = (long long)(nst >> 2);
    int threads = groups >= NTS ? NTS : (groups < 64 ? 64 : (int)groups);
    // forward: 32 B per amplitude per gate; backward: 2 passes of 32 B (all in shared memory; counted as the
    // same algorithmic bytes as the unfused kernels so that the classes stay comparable)
    qm_prof_work(QM_CLS_ENV, (double)batch * num_sweeps * n_gates * 96.0 * (double)nst);
    QM_LAUNCH(QM_CLS_ENV, st, k_sweeps_small<<<batch, threads, smem, st>>>(
        (const cplx*)targets, n_sites, (cplx*)gates, sites_dev, kinds_dev, n_gates, num_sweeps, (cplx*)envs, warm,
        (const cplx*)psis, (double*)overlaps));
    QM_CHECK_LAUNCH();
    return 0;
}

