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
// The vector FP64 pipe bounds the scalar passes (64 + 128 FP64 instructions per group, ~6 cycles per group and SM);
// mma.sync.m8n8k4.f64 does a group's gate in 1/4 instruction.  Fragment maps are chosen so that every shared-memory
// access is a 16-byte word per lane in a conflict-free pattern (8-byte accesses by component ran 4-way conflicted
// for gates on bits >= 3 and were SLOWER than the scalar code):
//  * apply, Y^T = X^T M^T: rows = the 8 groups, k = (re | im) of the 4 input amplitudes (two k-steps), columns =
//    (re, im) of the 4 outputs interleaved.  Lane (gr = lane>>2, tc = lane&3) loads amplitude tc of group G0+gr
//    (its re is the A operand of k-step 0, its im of k-step 1) and receives (re, im) of output tc of the same
//    group: one LDS.128, two DMMA, one STS.128 to the address it loaded from.
//  * environment, S[m][n] += sum_g (component m of tbar_g)(component n of c_g), components = (re_0..3, im_0..3), k =
//    4 groups per DMMA: lane loads amplitude gr&3 of group G0 + tc + 4(gr>>2) of both vectors (512 contiguous-per-
//    amplitude bytes per LDS.128) and swaps one double with lane^16, which gives it its operand for the DMMA of groups
//    G0..G0+3 and the one of groups G0+4..G0+7.
// The FP64 mma has a long latency (~200 cycles through a dependent chain): U independent chains per warp.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ int group_base4(int G, int q) { return ((G >> q) << (q + 2)) | (G & ((1 << q) - 1)); }

// B operands of the apply for the 4x4 matrix OP(G): b0 multiplies the re parts of the inputs, b1 the im parts
template <int OP>
__device__ __forceinline__ void gate_frag(const cplx* G, int lane, double& b0, double& b1) {
    const int gr = lane >> 2, tc = lane & 3;
    const cplx m = gate_elem<4, OP>(G, gr >> 1, tc);          // output gr>>1, input tc
    b0 = (gr & 1) ? m.y : m.x;
    b1 = (gr & 1) ? m.x : -m.y;
}

template <int U>
__device__ __forceinline__ void mma_apply(cplx* x, int G0, int q, double b0, double b1, int lane) {
    const int gr = lane >> 2, tc = lane & 3, stride = 1 << q;
    cplx* p[U];
    cplx v[U];
    double d0[U], d1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        p[u] = x + group_base4(G0 + 8 * u + gr, q) + tc * stride;
        v[u] = *p[u];
        d0[u] = 0.0;
        d1[u] = 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) dmma884(d0[u], d1[u], v[u].x, b0);
#pragma unroll
    for (int u = 0; u < U; u++) dmma884(d0[u], d1[u], v[u].y, b1);
#pragma unroll
    for (int u = 0; u < U; u++) *p[u] = mk(d0[u], d1[u]);
}

// warps w0 .. w0+nwork-1 of the CTA share the groups
template <int U>
__device__ __forceinline__ void small_apply_mma(cplx* x, int ngroups, int q, double b0, double b1, int wrel, int nwork,
                                                int lane) {
    for (int G0 = wrel * 8 * U; G0 < ngroups; G0 += nwork * 8 * U) mma_apply<U>(x, G0, q, b0, b1, lane);
}

__device__ __forceinline__ void small_apply_mma_any(cplx* x, int ngroups, int q, double b0, double b1, int wrel,
                                                    int nwork, int lane) {
    if (ngroups % (nwork * 64) == 0) small_apply_mma<8>(x, ngroups, q, b0, b1, wrel, nwork, lane);
    else if (ngroups % (nwork * 32) == 0) small_apply_mma<4>(x, ngroups, q, b0, b1, wrel, nwork, lane);
    else small_apply_mma<1>(x, ngroups, q, b0, b1, wrel, nwork, lane);
}

// Environment sums of the CURRENT vectors, S[m][n] += sum_g (component m of tbar_g)(component n of c_g); lane ends with
// S[gr][2tc], S[gr][2tc+1].  (The gate is taken out afterwards, E = S conj(G_old) in the polar warp, so that the update
// c <- G_old^H c is off the critical path: the other warps do it while warp 0 runs the polar.)
template <int U>
__device__ __forceinline__ void small_envS_mma(const cplx* c, const cplx* t, int ngroups, int q, int warp, int nw,
                                               int lane, double& e0, double& e1) {
    const int gr = lane >> 2, tc = lane & 3, stride = 1 << q;
    const bool lo = gr < 4;
    double f0[2 * U], f1[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; u++) { f0[u] = 0.0; f1[u] = 0.0; }
    for (int G0 = warp * 8 * U; G0 < ngroups; G0 += nw * 8 * U) {
        cplx zt[U], zy[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int b = group_base4(G0 + 8 * u + tc + 4 * (gr >> 2), q) + (gr & 3) * stride;
            zt[u] = t[b];
            zy[u] = c[b];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double rt = __shfl_xor_sync(0xffffffffu, lo ? zt[u].y : zt[u].x, 16);
            const double ry = __shfl_xor_sync(0xffffffffu, lo ? zy[u].y : zy[u].x, 16);
            dmma884(f0[2 * u], f1[2 * u], lo ? zt[u].x : rt, lo ? zy[u].x : ry);               // groups G0+8u .. +3
            dmma884(f0[2 * u + 1], f1[2 * u + 1], lo ? rt : zt[u].y, lo ? ry : zy[u].y);       // groups G0+8u+4 .. +7
        }
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
               int defer_mode, long long* __restrict__ dbg) {
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
                small_apply_mma_any(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
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
            double gh0 = 0.0, gh1 = 0.0;                      // fragments of G_old^H for the deferred c update
            if (mma) {
                double e0 = 0.0, e1 = 0.0;
                if (mma4) small_envS_mma<4>(c, t, nst >> 2, q, warp, nw, lane, e0, e1);
                else small_envS_mma<1>(c, t, nst >> 2, q, warp, nw, lane, e0, e1);
                wsum[warp][2 * lane] = e0;
                wsum[warp][2 * lane + 1] = e1;
                gate_frag<1>(G, lane, gh0, gh1);
                if (defer_mode == 0) {                       // same warp -> same groups as its part of the S pass
                    __syncwarp();
                    if (mma4) small_apply_mma<4>(c, nst >> 2, q, gh0, gh1, warp, nw, lane);
                    else small_apply_mma<1>(c, nst >> 2, q, gh0, gh1, warp, nw, lane);
                }
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
                    // S is the real 8x8 product of (re, im) components: E'[o][b'] = sum tbar_o c_b' with the OLD gate
                    // still inside c; the environment of the gate is E[o][b] = sum_b' E'[o][b'] conj(G_old[b'][b])
                    if (lane < 16) {
                        const int o = lane >> 2, b = lane & 3;
                        pol_scratch[lane] = mk(ssum64[o * 8 + b] - ssum64[(4 + o) * 8 + 4 + b],
                                               ssum64[o * 8 + 4 + b] + ssum64[(4 + o) * 8 + b]);
                    }
                    __syncwarp();
                    if (lane < 16) {
                        const int o = lane >> 2, b = lane & 3;
                        const cplx p0 = cmulc(pol_scratch[o * 4 + 0], G[0 * 4 + b]), p1 = cmulc(pol_scratch[o * 4 + 1], G[1 * 4 + b]);
                        const cplx p2 = cmulc(pol_scratch[o * 4 + 2], G[2 * 4 + b]), p3 = cmulc(pol_scratch[o * 4 + 3], G[3 * 4 + b]);
                        Es[lane] = cadd(cadd(p0, p1), cadd(p2, p3));
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
            } else if (mma && defer_mode == 1) {
                small_apply_mma_any(c, nst >> 2, q, gh0, gh1, warp - 1, nw - 1, lane);    // c <- G_old^H c, under the polar
            } else if (mma && defer_mode == 2) {
                // ... by the warps that do not share warp 0's scheduler (warp % 4 != 0), so that the polar's chain of
                // dependent FP64 instructions does not queue behind their DMMAs
                if (nw < 4) small_apply_mma_any(c, nst >> 2, q, gh0, gh1, warp - 1, nw - 1, lane);
                else if (warp & 3) small_apply_mma_any(c, nst >> 2, q, gh0, gh1, (warp >> 2) * 3 + (warp & 3) - 1, (nw >> 2) * 3, lane);
            }
            const long long c2 = dbg ? clock64() : 0;
            __syncthreads();
            if (mma) {
                double a0, a1;
                gate_frag<2>(G, lane, a0, a1);
                small_apply_mma_any(t, nst >> 2, q, a0, a1, warp, nw, lane);
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
                small_apply_mma_any(c, nst >> 2, gq[k], a0, a1, warp, nw, lane);
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
    // DMMA passes: measured faster than the scalar ones up to 10 qubits, slower at 12 (the FP64 tensor pipe of this
    // chip has the vector pipe's peak and a long latency; profiles/sweeps_small_phases_r02*.log); QM_SMALL_MMA=0/1 forces
    static const int mma_env = getenv("QM_SMALL_MMA") ? atoi(getenv("QM_SMALL_MMA")) : -1;
    const int allow_mma = mma_env >= 0 ? mma_env : (n_sites <= 10);
    static const int defer_mode = getenv("QM_SMALL_DEFER") ? atoi(getenv("QM_SMALL_DEFER")) : 2;
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
        (const cplx*)psis, (double*)overlaps, allow_mma, defer_mode, dbg));
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

