// Small registers: all optimisation sweeps of a state in one launch (SURVEY 8a rows A8 + A9 for states whose two
// dense vectors fit one SM's shared memory).  Reference: qmprs/synthesis/mps_encoding/sequential.py:443-505, 509-541.
// Own translation unit: see polar.cuh.
#include "common.cuh"
#include "polar.cuh"
#include "qmprs_b200.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

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

// ---- FP64 tensor-pipe variants of the two passes for two-qubit gates (8 groups per warp-step) ------------------
// A 4x4 complex gate acting on a group is the real 8x8 matrix [[Re M, -Im M], [Im M, Re M]] acting on the group's
// (re_0..re_3, im_0..im_3): one mma.sync.m8n8k4.f64 pair applies it to 8 groups.  Lane (gr = lane>>2, tc = lane&3)
// loads amplitude tc of group G0+gr as ONE 16-byte word (its re feeds k-step 0, its im k-step 1) and receives
// component gr of the results of groups G0+2tc, G0+2tc+1.  The vector FP64 pipe issues ~1 warp-instruction per cycle
// per SM, which bounded the scalar passes (64 + 128 FP64 instructions per group); here a group costs 1/4 DMMA.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ int group_base4(int G, int q) { return ((G >> q) << (q + 2)) | (G & ((1 << q) - 1)); }

template <int OP>
__device__ __forceinline__ void gate_frag(const cplx* G, int lane, double& a0, double& a1) {
    const int gr = lane >> 2, tc = lane & 3;
    const cplx m = gate_elem<4, OP>(G, gr & 3, tc);
    a0 = gr < 4 ? m.x : m.y;
    a1 = gr < 4 ? -m.y : m.x;
}

// y = M x for U x 8 groups starting at G0, written back in place.  U independent DMMA chains per warp: the FP64
// mma has a long latency (~200 cycles measured through this loop), a single chain per warp ran SLOWER than the scalar
// passes.
template <int U>
__device__ __forceinline__ void mma_apply(cplx* x, int G0, int q, double a0, double a1, int lane) {
    const int gr = lane >> 2, tc = lane & 3, stride = 1 << q;
    cplx v[U];
    double d0[U], d1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        v[u] = x[group_base4(G0 + 8 * u + gr, q) + tc * stride];
        d0[u] = 0.0;
        d1[u] = 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) dmma884(d0[u], d1[u], a0, v[u].x);
#pragma unroll
    for (int u = 0; u < U; u++) dmma884(d0[u], d1[u], a1, v[u].y);
#pragma unroll
    for (int u = 0; u < U; u++) {
        double* p0 = (double*)&x[group_base4(G0 + 8 * u + 2 * tc, q) + (gr & 3) * stride] + (gr >> 2);
        double* p1 = (double*)&x[group_base4(G0 + 8 * u + 2 * tc + 1, q) + (gr & 3) * stride] + (gr >> 2);
        *p0 = d0[u];
        *p1 = d1[u];
    }
}

template <int U>
__device__ __forceinline__ void small_apply_mma(cplx* x, int ngroups, int q, double a0, double a1, int warp, int nw,
                                                int lane) {
    for (int G0 = warp * 8 * U; G0 < ngroups; G0 += nw * 8 * U) mma_apply<U>(x, G0, q, a0, a1, lane);
}

// c <- G^H c, then S[m][n] += sum_groups (component m of tbar) (component n of c_new); lane holds S[gr][2tc], S[gr][2tc+1]
template <int U>
__device__ __forceinline__ void small_env_mma(cplx* c, const cplx* t, int ngroups, int q, double a0, double a1, int warp,
                                              int nw, int lane, double& e0, double& e1) {
    const int gr = lane >> 2, tc = lane & 3, stride = 1 << q;
    double f0[2 * U], f1[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; u++) { f0[u] = 0.0; f1[u] = 0.0; }
    for (int G0 = warp * 8 * U; G0 < ngroups; G0 += nw * 8 * U) {
        mma_apply<U>(c, G0, q, a0, a1, lane);
        __syncwarp();
        double ta[2 * U], yb[2 * U];
#pragma unroll
        for (int u = 0; u < 2 * U; u++) {
            const int b = group_base4(G0 + 4 * u + tc, q) + (gr & 3) * stride;
            ta[u] = ((const double*)&t[b])[gr >> 2];
            yb[u] = ((const double*)&c[b])[gr >> 2];
        }
#pragma unroll
        for (int u = 0; u < 2 * U; u++) dmma884(f0[u], f1[u], ta[u], yb[u]);
    }
#pragma unroll
    for (int h = U; h >= 1; h >>= 1)                     // fixed tree over the independent accumulators
#pragma unroll
        for (int u = 0; u < h; u++) { f0[u] += f0[u + h]; f1[u] += f1[u + h]; }
    e0 = f0[0];
    e1 = f1[0];
}

__global__ void __launch_bounds__(NTS, 1)
k_sweeps_small(const cplx* __restrict__ targets, int nbits, cplx* __restrict__ gates_g, const int* __restrict__ sites,
               const int* __restrict__ kinds, int n_gates, int num_sweeps, cplx* __restrict__ envs_g, int warm,
               const cplx* __restrict__ psis, double* __restrict__ overlaps, int allow_mma,
               long long* __restrict__ dbg) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    const int nst = 1 << nbits;
    cplx* c = (cplx*)sw_smem;
    cplx* t = c + nst;
    cplx* g = t + nst;                                   // [n_gates][16]
    // warm start of each gate's 4x4 Jacobi polar: right singular vectors of the previous sweep (as qm_sweep_stored)
    cplx* vw = warm ? g + (long long)n_gates * 16 : nullptr;
    int* gq = (int*)(g + (long long)n_gates * 16 * (warm ? 2 : 1));     // lowest bit of each gate
    int* gd = gq + n_gates;                              // dimension (2 or 4)
    __shared__ double wsum[NTS / 32][64];
    __shared__ double ssum64[64];
    __shared__ cplx Es[16];
    __shared__ cplx pol_scratch[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const bool use_mma = allow_mma && (nst >> 2) >= 8 * nw;              // every warp has whole 8-group steps
    const bool mma4 = (nst >> 2) >= 32 * nw;                             // ... and whole steps of 4 x 8 groups
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
            if (gd[k] == 4 && use_mma) {
                double a0, a1;
                gate_frag<0>(g + k * 16, lane, a0, a1);
                if (mma4) small_apply_mma<4>(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
                else small_apply_mma<1>(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
            } else if (gd[k] == 4) small_apply<4, 0>(c, nst, gq[k], g + k * 16);
            else small_apply<2, 0>(c, nst, gq[k], g + k * 16);
            __syncthreads();
        }
        for (int i = tid; i < nst; i += blockDim.x) t[i] = cconj(target[i]);
        __syncthreads();
        for (int k = n_gates - 1; k >= 0; k--) {
            const int d = gd[k], q = gq[k];
            cplx* G = g + k * 16;
            const long long c0 = dbg ? clock64() : 0;
            const bool mma = d == 4 && use_mma;
            if (mma) {
                double a0, a1, e0 = 0.0, e1 = 0.0;
                gate_frag<1>(G, lane, a0, a1);
                if (mma4) small_env_mma<4>(c, t, nst >> 2, q, a0, a1, warp, nw, lane, e0, e1);
                else small_env_mma<1>(c, t, nst >> 2, q, a0, a1, warp, nw, lane, e0, e1);
                wsum[warp][2 * lane] = e0;
                wsum[warp][2 * lane + 1] = e1;
            } else {
                double acc[32];
#pragma unroll
                for (int i = 0; i < 32; i++) acc[i] = 0.0;
                if (d == 4) small_env<4>(c, t, nst, q, G, acc);
                else small_env<2>(c, t, nst, q, G, acc);
                wsum[warp][lane] = warp_reduce32(acc, lane);
            }
            __syncthreads();
            const long long c1 = dbg ? clock64() : 0;
            if (warp == 0) {
                if (mma) {
                    double v0[NTS / 32], v1[NTS / 32];                      // fixed pairwise tree over the warps
#pragma unroll
                    for (int w = 0; w < NTS / 32; w++) {
                        v0[w] = w < nw ? wsum[w][2 * lane] : 0.0;
                        v1[w] = w < nw ? wsum[w][2 * lane + 1] : 0.0;
                    }
#pragma unroll
                    for (int h = NTS / 64; h >= 1; h >>= 1)
#pragma unroll
                        for (int w = 0; w < h; w++) {
                            v0[w] += v0[w + h];
                            v1[w] += v1[w + h];
                        }
                    ssum64[2 * lane] = v0[0];
                    ssum64[2 * lane + 1] = v1[0];
                    __syncwarp();
                    // S is the real 8x8 product of (re, im) components: E[o][b] = sum tbar_o c_b
                    if (lane < 16) {
                        const int o = lane >> 2, b = lane & 3;
                        Es[lane] = mk(ssum64[o * 8 + b] - ssum64[(4 + o) * 8 + 4 + b],
                                      ssum64[o * 8 + 4 + b] + ssum64[(4 + o) * 8 + b]);
                    }
                } else {
                    double ssum = 0.0;
                    for (int w = 0; w < nw; w++) ssum += wsum[w][lane];     // fixed order
                    // lane holds one double of E: (re, im) of entry lane/2
                    const double other = __shfl_xor_sync(0xffffffffu, ssum, 1);
                    if ((lane & 1) == 0 && (lane >> 1) < d * d) Es[lane >> 1] = mk(ssum, other);
                }
                __syncwarp();
                const int pr = polar_conj_warp(Es, d, G, pol_scratch, warm ? vw + k * 16 : nullptr,
                                               warm ? vw + k * 16 : nullptr);
                if (dbg && tid == 0) {
                    dbg[blockIdx.x * 5 + 3] += pr < 0 ? -pr : pr;
                    dbg[blockIdx.x * 5 + 4] += pr < 0;
                }
                // the rank-deficient branch of the polar returns early in 31 lanes while lane 0 finishes the
                // single-thread completion: reconverge before the block barrier
                __syncwarp();
                if (envs && sweep == num_sweeps - 1 && lane < d * d) envs[k * 16 + lane] = Es[lane];
            }
            const long long c2 = dbg ? clock64() : 0;
            __syncthreads();
            if (mma) {
                double a0, a1;
                gate_frag<2>(G, lane, a0, a1);
                if (mma4) small_apply_mma<4>(t, nst >> 2, q, a0, a1, warp, nw, lane);
                else small_apply_mma<1>(t, nst >> 2, q, a0, a1, warp, nw, lane);
            } else if (d == 4) small_apply<4, 2>(t, nst, q, G);
            else small_apply<2, 2>(t, nst, q, G);
            __syncthreads();
            if (dbg && tid == 0) {                 // QM_SMALL_DEBUG: cycles of env pass / reduce + polar / tbar update
                const long long c3 = clock64();
                dbg[blockIdx.x * 5] += c1 - c0;
                dbg[blockIdx.x * 5 + 1] += c2 - c1;
                dbg[blockIdx.x * 5 + 2] += c3 - c2;
            }
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
            if (gd[k] == 4 && use_mma) {
                double a0, a1;
                gate_frag<0>(g + k * 16, lane, a0, a1);
                if (mma4) small_apply_mma<4>(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
                else small_apply_mma<1>(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
            } else if (gd[k] == 4) small_apply<4, 0>(c, nst, gq[k], g + k * 16);
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
    const int warm = smem + (size_t)n_gates * 16 * sizeof(cplx) <= 212 * 1024;      // room for the polar warm starts
    if (warm) smem += (size_t)n_gates * 16 * sizeof(cplx);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        QM_CUDA(cudaFuncSetAttribute(k_sweeps_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    static const int allow_mma = getenv("QM_SMALL_MMA") ? atoi(getenv("QM_SMALL_MMA")) : 1;
    static const int debug = getenv("QM_SMALL_DEBUG") ? atoi(getenv("QM_SMALL_DEBUG")) : 0;
    long long* dbg = nullptr;
    if (debug) {
        QM_CUDA(cudaMalloc(&dbg, (size_t)batch * 5 * sizeof(long long)));
        QM_CUDA(cudaMemsetAsync(dbg, 0, (size_t)batch * 5 * sizeof(long long), st));
    }
    long long groups = (long long)(nst >> 2);
    int threads = groups >= NTS ? NTS : (groups < 64 ? 64 : (int)groups);
    // forward: 32 B per amplitude per gate; backward: 2 passes of 32 B (all in shared memory; counted as the
    // same algorithmic bytes as the unfused kernels so that the classes stay comparable)
    qm_prof_work(QM_CLS_ENV, (double)batch * num_sweeps * n_gates * 96.0 * (double)nst);
    QM_LAUNCH(QM_CLS_ENV, st, k_sweeps_small<<<batch, threads, smem, st>>>(
        (const cplx*)targets, n_sites, (cplx*)gates, sites_dev, kinds_dev, n_gates, num_sweeps, (cplx*)envs, warm,
        (const cplx*)psis, (double*)overlaps, allow_mma, dbg));
    QM_CHECK_LAUNCH();
    if (dbg) {
        std::vector<long long> h((size_t)batch * 5);
        QM_CUDA(cudaStreamSynchronize(st));
        QM_CUDA(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        QM_CUDA(cudaFree(dbg));
        double ph[5] = {0, 0, 0, 0, 0};
        for (int b = 0; b < batch; b++)
            for (int j = 0; j < 5; j++) ph[j] += (double)h[(size_t)b * 5 + j];
        const double steps = (double)batch * num_sweeps * n_gates;
        fprintf(stderr, "[qm_sweeps_small] cycles per gate-step: env pass %.0f, reduce + polar %.0f, tbar update %.0f; polar rounds %.2f, "
                "rank-deficient %.4f\n", ph[0] / steps, ph[1] / steps, ph[2] / steps, ph[3] / steps, ph[4] / steps);
    }
    return 0;
}

