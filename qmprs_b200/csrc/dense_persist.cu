// Optimisation sweeps of a large register in ONE persistent launch (SURVEY 8a rows A8 + A9).
//
// Reference: qmprs/synthesis/mps_encoding/sequential.py:509-541 (_optimize_unitary_layers) driving
//   :215-292, :443-447   dense circuit state from the current gates          (forward pass)
//   :452-505             per gate, last applied first:  E = tbar . c_k ;  G_new = conj(polar(E)) ;
//                        tbar <- G_new^T tbar                                 (backward pass)
//
// Round 1 ran one kernel per gate (k_gate2 forward, k_env_fused backward: 30 us per gate-step at 20 qubits where
// the traffic needs ~8 us).  The chain is strictly serial -- gate k's environment needs the gate updated in step
// k+1 -- so what bounds a step is the latency between "last partial sum written" and "new gate visible to
// every SM": a ticket, the last CTA re-reading 512 partials, a one-warp polar, kernel completion, the next
// launch.  Here the whole optimisation (all sweeps, forward and backward passes) is one cooperative launch with
// one CTA per SM and ONE grid barrier per gate-step:
//   * every CTA streams its share of the amplitudes (same fused pass as k_env_fused: read tbar, apply the
//     pending gate, write tbar, read the stored c_k, accumulate the 4x4 environment), writes ONE partial;
//   * grid barrier (release add / acquire spin on a monotonically increasing counter: no reset, no second phase);
//   * EVERY CTA then sums the gridDim.x partials in the same fixed order and runs the same polar update --
//     redundant, identical to the bit, and it removes the broadcast: the new gate is already in every CTA's
//     shared memory when the next step starts.  CTA 0 alone writes the gate (and environment / warm start) back.
// Partials are double-buffered by step parity (a CTA may start step k-1 while another still reads step k's).
// Intermediate circuit states c_k are kept in HBM as in round 1 ((M+1) x 2^N amplitudes).
// Own translation unit: see polar.cuh.
#include "common.cuh"
#include "polar.cuh"
#include "qmprs_b200.h"
#include <cstdlib>
#include <cstdio>

namespace {

constexpr int NTP = 256;
constexpr int MAXBLK = 148 * 2;

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch) {
    __syncthreads();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();                                   // publish this CTA's writes (cumulative over the CTA barrier)
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - epoch) < 0);
        __threadfence();
    }
    __syncthreads();
}

// x_out <- M x_in on bits (q+1, q) [D = 4] or bit q [D = 2]; M row-major in shared memory
template <int D>
__device__ __forceinline__ void stream_gate(const cplx* xin, cplx* xout, int nbits, int q, const cplx* M) {
    constexpr int K = (D == 4) ? 2 : 1;
    const long long ngroups = 1LL << (nbits - K);
    const long long stride = 1LL << q, lowmask = stride - 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ngroups;
         t += (long long)gridDim.x * blockDim.x) {
        const long long base = ((t >> q) << (q + K)) | (t & lowmask);
        cplx v[D], y[D];
#pragma unroll
        for (int a = 0; a < D; a++) v[a] = xin[base + a * stride];
#pragma unroll
        for (int a = 0; a < D; a++) {
            cplx s = cmul(M[a * D + 0], v[0]);
#pragma unroll
            for (int b = 1; b < D; b++) cfma(s, M[a * D + b], v[b]);
            y[a] = s;
        }
#pragma unroll
        for (int a = 0; a < D; a++) xout[base + a * stride] = y[a];
    }
}

// One fused backward pass over this CTA's share (body of round 1's k_env_fused).  Window [q0, q0+NU): the current
// gate (dimension CD) acts on its top bits, the pending gate (dimension PD, 0 = none) on the bottom ones;
// P = pending gate TRANSPOSED (tbar'[b] = sum_o G[o][b] tbar[o]) in shared memory.  acc[2(o CD + b)] += tbar' . c
template <int NU, int CD, int PD>
__device__ __forceinline__ void env_pass(cplx* tbar, const cplx* c, int nbits, int q0, const cplx* P, double* acc) {
    constexpr int GSZ = 1 << NU;
    constexpr int CSH = NU - (CD == 4 ? 2 : 1);
    constexpr int NLOW = 1 << CSH;
    constexpr int PDD = PD > 0 ? PD : 1;
    const long long ngroups = 1LL << (nbits - NU);
    const long long stride = 1LL << q0, lowmask = stride - 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ngroups;
         t += (long long)gridDim.x * blockDim.x) {
        const long long base = ((t >> q0) << (q0 + NU)) | (t & lowmask);
        cplx tv[GSZ], cv[GSZ];
#pragma unroll
        for (int j = 0; j < GSZ; j++) cv[j] = __ldcs(c + base + j * stride);      // streamed once: do not keep in L2
#pragma unroll
        for (int j = 0; j < GSZ; j++) tv[j] = tbar[base + j * stride];
        if (PD > 0) {
#pragma unroll
            for (int h = 0; h < GSZ / PDD; h++) {
                cplx y[PDD];
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
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct Smem {
    double wsum[NTP / 32][32];
    double Ep[NTP];
    cplx Es[16];
    cplx Gnew[16];          // gate produced by the last polar update (compact d x d, row-major)
    cplx M[16];             // operand matrix of the current streaming pass
    cplx scratch[32];
};

// CTA partial of the 32 accumulators -> partials[blockIdx.x][32]
__device__ __forceinline__ void cta_partial(double* acc, Smem& sm, double* partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double mine = warp_reduce32(acc, lane);
    sm.wsum[warp][lane] = mine;
    __syncthreads();
    if (warp == 0) {
        double ssum = 0.0;
#pragma unroll
        for (int w = 0; w < NTP / 32; w++) ssum += sm.wsum[w][lane];
        partials[(long long)blockIdx.x * 32 + lane] = ssum;
    }
}

// every CTA: fixed-order sum of all partials, polar update into sm.Gnew; CTA 0 writes the results back
template <int CD>
__device__ __forceinline__ void reduce_polar(const double* partials, Smem& sm, cplx* gate_out, cplx* env_out,
                                             const cplx* vw_in, cplx* vw_out) {
    {
        const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
        // all of a thread's partials in flight at once (one L2 round trip instead of three), summed in the same order
        constexpr int NLD = (MAXBLK + NTP / 32 - 1) / (NTP / 32);
        double v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const unsigned int b = sl + k * (NTP / 32);
            v[k] = b < gridDim.x ? __ldcg(partials + (long long)b * 32 + e) : 0.0;
        }
        double ssum = 0.0;
#pragma unroll
        for (int k = 0; k < NLD; k++) ssum += v[k];
        sm.Ep[threadIdx.x] = ssum;
        __syncthreads();
        if (threadIdx.x < 32) {
            double tsum = 0.0;
#pragma unroll
            for (int k = 0; k < NTP / 32; k++) tsum += sm.Ep[k * 32 + threadIdx.x];
            sm.Ep[threadIdx.x] = tsum;
        }
        __syncthreads();
        if (threadIdx.x < CD * CD) sm.Es[threadIdx.x] = mk(sm.Ep[2 * threadIdx.x], sm.Ep[2 * threadIdx.x + 1]);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        polar_conj_warp(sm.Es, CD, sm.Gnew, sm.scratch, vw_in, blockIdx.x == 0 ? vw_out : nullptr);
        __syncwarp();
        if (blockIdx.x == 0 && threadIdx.x < CD * CD) {
            gate_out[threadIdx.x] = sm.Gnew[threadIdx.x];
            if (env_out) env_out[threadIdx.x] = sm.Es[threadIdx.x];
        }
    }
    __syncthreads();
}

// lowest bit of a gate on `site` of kind 1 / 2 in an nbits register
__device__ __forceinline__ int low_bit(int nbits, int site, int kind) { return kind == 2 ? nbits - 2 - site : nbits - 1 - site; }

__global__ void __launch_bounds__(NTP, 1)
k_sweeps_persist(cplx* cs, cplx* tbar, const cplx* __restrict__ target, int nbits, cplx* gates,
                 const int* __restrict__ sites, const int* __restrict__ kinds, int n_gates, int num_sweeps,
                 double* partials, unsigned int* bar, cplx* vwarm /* [2][n_gates][16] */, cplx* envs, int prefetch,
                 unsigned long long* dbg) {
    // (no __restrict__ on buffers this kernel both writes and reads across grid barriers: their loads must stay
    // coherent, never ld.global.nc)
    __shared__ Smem sm;
    const long long n = 1LL << nbits;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    unsigned int epoch = 0;
    int par = 0;
    for (int sweep = 0; sweep < num_sweeps; sweep++) {
        const cplx* vw_in = vwarm + (long long)(sweep & 1) * n_gates * 16;
        cplx* vw_out = vwarm + (long long)((sweep + 1) & 1) * n_gates * 16;
        // ---- forward pass: cs[0] = |0..0>, cs[k+1] = G_k cs[k]; tbar = conj(target) rides along ----
        for (long long i = gtid; i < n; i += gthreads) {
            cs[i] = mk(i == 0 ? 1.0 : 0.0, 0.0);
            const cplx t = target[i];
            tbar[i] = mk(t.x, -t.y);
        }
        grid_barrier(bar, epoch);
        const bool timing = dbg && blockIdx.x == 0 && threadIdx.x == 0;       // QM_PERSIST_DEBUG: ns per phase, CTA 0
        unsigned long long t0 = timing ? gtimer() : 0;
        for (int g = 0; g < n_gates; g++) {
            const int kind = kinds[g], site = sites[g];
            if (threadIdx.x < 16) sm.M[threadIdx.x] = gates[(long long)g * 16 + threadIdx.x];
            __syncthreads();
            const cplx* xin = cs + (long long)g * n;
            cplx* xout = cs + (long long)(g + 1) * n;
            if (kind == 2) stream_gate<4>(xin, xout, nbits, low_bit(nbits, site, 2), sm.M);
            else stream_gate<2>(xin, xout, nbits, low_bit(nbits, site, 1), sm.M);
            grid_barrier(bar, epoch);
        }
        if (timing) { const unsigned long long t1 = gtimer(); dbg[0] += t1 - t0; t0 = t1; }
        // ---- backward pass ----
        for (int g = n_gates - 1; g >= 0; g--) {
            const int ck = kinds[g];
            const int cb = low_bit(nbits, sites[g], ck);
            const cplx* c = cs + (long long)g * n;
            int pk = 0, pb = 0, ptop = 0;
            if (g + 1 < n_gates) {
                pk = kinds[g + 1];
                pb = low_bit(nbits, sites[g + 1], pk);
                ptop = pb + (pk == 2 ? 1 : 0);
            }
            // which fused variant (same cases as qm_sweep_stored)
            int variant = 0;                               // 0: no fusion
            if (pk) {
                if (ck == 2 && pk == 2 && ptop == cb) variant = 1;            // <3,4,4>
                else if (ck == 2 && pk == 1 && pb == cb) variant = 2;         // <2,4,2>
                else if (ck == 1 && pk == 2 && ptop == cb - 1) variant = 3;   // <3,2,4>
                else if (ck == 1 && pk == 1 && pb == cb - 1) variant = 4;     // <2,2,2>
            }
            if (pk) {
                // pending gate = sm.Gnew (compact d x d).  Fused: transposed into M;  unfused: one extra pass
                const int pd = pk == 2 ? 4 : 2;
                if (threadIdx.x < pd * pd) {
                    const int a = threadIdx.x / pd, b = threadIdx.x % pd;
                    sm.M[threadIdx.x] = sm.Gnew[b * pd + a];
                }
                __syncthreads();
                if (variant == 0) {
                    if (pk == 2) stream_gate<4>(tbar, tbar, nbits, pb, sm.M);
                    else stream_gate<2>(tbar, tbar, nbits, pb, sm.M);
                    grid_barrier(bar, epoch);
                }
            }
            double acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.0;
            switch (variant) {
                case 1: env_pass<3, 4, 4>(tbar, c, nbits, pb, sm.M, acc); break;
                case 2: env_pass<2, 4, 2>(tbar, c, nbits, pb, sm.M, acc); break;
                case 3: env_pass<3, 2, 4>(tbar, c, nbits, pb, sm.M, acc); break;
                case 4: env_pass<2, 2, 2>(tbar, c, nbits, pb, sm.M, acc); break;
                default:
                    if (ck == 2) env_pass<2, 4, 0>(tbar, c, nbits, cb, sm.M, acc);
                    else env_pass<1, 2, 0>(tbar, c, nbits, cb, sm.M, acc);
            }
            if (timing) { const unsigned long long t1 = gtimer(); dbg[1] += t1 - t0; t0 = t1; }
            double* part = partials + (long long)par * MAXBLK * 32;
            cta_partial(acc, sm, part);
            if (prefetch && g > 0) {
                // HBM idles during the barrier + reduction + polar that follow: pull the next step's stored circuit
                // state (read once, last touched a whole forward pass ago) into L2 meanwhile
                const char* nxt = (const char*)(cs + (long long)(g - 1) * n);
                const long long lines = (n * (long long)sizeof(cplx)) >> 7;
                for (long long i = gtid; i < lines; i += gthreads)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (i << 7)));
            }
            grid_barrier(bar, epoch);
            if (timing) { const unsigned long long t1 = gtimer(); dbg[2] += t1 - t0; t0 = t1; }
            cplx* G = gates + (long long)g * 16;
            cplx* env = envs ? envs + (long long)g * 16 : nullptr;
            if (ck == 2) reduce_polar<4>(part, sm, G, env, vw_in + (long long)g * 16, vw_out + (long long)g * 16);
            else reduce_polar<2>(part, sm, G, env, vw_in + (long long)g * 16, vw_out + (long long)g * 16);
            par ^= 1;
            if (timing) { const unsigned long long t1 = gtimer(); dbg[3] += t1 - t0; t0 = t1; }
        }
        grid_barrier(bar, epoch);          // gates of this sweep (written by CTA 0) visible to every CTA's forward pass
    }
}

}  // namespace

extern "C" long long qm_sweeps_persist_work_bytes(int n_gates) {
    return 256 + 2LL * MAXBLK * 32 * sizeof(double) + 2LL * n_gates * 16 * sizeof(cplx);
}

// All `num_sweeps` optimisation sweeps in one cooperative launch.  cs: (n_gates+1) x 2^N scratch for the stored
// circuit states; tbar: 2^N scratch; target: the dense target (not conjugated); gates: [n_gates][16] updated in
// place; sites_dev / kinds_dev: DEVICE int[n_gates]; work: qm_sweeps_persist_work_bytes(n_gates) bytes;
// envs: optional [n_gates][16], environments of the last sweep.  Returns -3 if the device cannot co-schedule
// the grid (the caller falls back to qm_circuit_states + qm_sweep_stored).
extern "C" int qm_sweeps_persist(void* cs, void* tbar, const void* target, int n_sites, void* gates,
                                 const int* sites_dev, const int* kinds_dev, int n_gates, int num_sweeps, void* work,
                                 void* envs, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_gates <= 0 || num_sweeps <= 0) return 0;
    static int n_sm = 0, per_sm = 0;
    if (!n_sm) {
        int dev = 0;
        QM_CUDA(cudaGetDevice(&dev));
        QM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweeps_persist, NTP, 0));
    }
    if (per_sm < 1) return -3;
    // one CTA per SM; fewer when the register is so small that a CTA would have nothing to stream
    long long want = (1LL << n_sites) / 8 / NTP;
    int grid = n_sm;
    if (want < grid) grid = want < 1 ? 1 : (int)want;
    if (grid > MAXBLK) grid = MAXBLK;
    unsigned int* bar = (unsigned int*)work;
    double* partials = (double*)((char*)work + 256);
    cplx* vwarm = (cplx*)((char*)work + 256 + 2LL * MAXBLK * 32 * sizeof(double));
    QM_CUDA(cudaMemsetAsync(bar, 0, 256, st));
    QM_CUDA(cudaMemsetAsync(vwarm, 0, 2LL * n_gates * 16 * sizeof(cplx), st));
    cplx* cs_ = (cplx*)cs; cplx* tbar_ = (cplx*)tbar; const cplx* target_ = (const cplx*)target;
    cplx* gates_ = (cplx*)gates; cplx* envs_ = (cplx*)envs;
    static int prefetch = getenv("QM_PERSIST_PREFETCH") ? atoi(getenv("QM_PERSIST_PREFETCH")) : 1;
    static const int debug = getenv("QM_PERSIST_DEBUG") ? atoi(getenv("QM_PERSIST_DEBUG")) : 0;
    unsigned long long* dbg = nullptr;
    if (debug) {
        QM_CUDA(cudaMalloc(&dbg, 4 * sizeof(unsigned long long)));
        QM_CUDA(cudaMemsetAsync(dbg, 0, 4 * sizeof(unsigned long long), st));
    }
    void* args[] = {&cs_, &tbar_, &target_, &n_sites, &gates_, &sites_dev, &kinds_dev, &n_gates, &num_sweeps,
                    &partials, &bar, &vwarm, &envs_, &prefetch, &dbg};
    // forward: read c_k, write c_{k+1} (32 B per amplitude and gate); backward: tbar read + write, c_k read (48 B)
    qm_prof_work(QM_CLS_ENV, 80.0 * (double)(1LL << n_sites) * n_gates * num_sweeps);
    QM_LAUNCH(QM_CLS_ENV, st, cudaLaunchCooperativeKernel((void*)k_sweeps_persist, dim3(grid), dim3(NTP), args, 0, st));
    QM_CHECK_LAUNCH();
    if (dbg) {
        unsigned long long h[4];
        QM_CUDA(cudaStreamSynchronize(st));
        QM_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        QM_CUDA(cudaFree(dbg));
        const double steps = (double)n_gates * num_sweeps;
        fprintf(stderr, "[qm_sweeps_persist] us per gate-step (CTA 0): forward %.2f, backward pass %.2f, partial + barrier "
                        "%.2f, reduce + polar %.2f\n", h[0] / steps / 1e3, h[1] / steps / 1e3, h[2] / steps / 1e3,
                h[3] / steps / 1e3);
    }
    return 0;
}
