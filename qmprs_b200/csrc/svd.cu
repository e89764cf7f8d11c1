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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "qmprs_b200.h"

namespace {

constexpr int BSZ = 16;    // rows per block
constexpr int PMAX = 32;   // rows per pair (Gram order)
constexpr int TC = 32;     // tile columns
constexpr int NT = 256;

struct Geom {
    int nv, len, nvp, nbp, single, nrows, npairs, rounds, ext;
    long long ldw;
};

// ext = 1: Wext = [W | I] accumulates the rotations (both factors from the iteration).  ext = 0 (multi-block
// only): W alone is rotated and the second factor is recovered at the end by one ZGEMM against the input
// (QM_SVD_BACKMULT, see svd_impl): a third less tensor work per sweep on square matrices.
Geom make_geom(int m, int n, int backmult = 0) {
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
    g.ext = (backmult && !g.single) ? 0 : 1;
    g.ldw = (long long)g.len + (g.ext ? g.nvp : 0);
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

// Which block pair a CTA works on.  mode 0: slot p of round r of the circle tournament over the n blocks
// [off_a, off_a + n);  mode 1: cross pairs between two disjoint groups of n blocks, (off_a + p, off_b + (p + r) % n).
// The grouped schedule (svd_impl) builds a sweep from two half-size tournaments followed by two rounds of
// quarter x quarter cross products, so that two independent pair streams exist at every moment.
struct PairSpec { int mode, r, n, off_a, off_b; };
__device__ __forceinline__ void get_pair(const PairSpec& ps, int p, int& bi, int& bj) {
    if (ps.mode == 0) {
        circle_pair(ps.r, p, ps.n, bi, bj);
        bi += ps.off_a; bj += ps.off_a;
    } else {
        bi = ps.off_a + p;
        bj = ps.off_b + (p + ps.r) % ps.n;
    }
}

// ---------------------------------------------------------------------------------
// (1) Gram matrices.  grid = (nchunks, npairs).  Every CTA writes the partial Gram matrix of its column
// chunk as a compact slab; k_eig sums the slabs in fixed order (no atomics: reproducible).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_gram(const cplx* __restrict__ W, long long ldw, int len, int chunk, int round, int nbp, int single,
       int nrows, double* __restrict__ G, const int* __restrict__ done) {
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
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
    // no atomics: run-to-run reproducible.  Column slices of an entry are combined in fixed order through shared
    // memory, and the CTA writes its own compact slab [E][2]; k_eig sums the slabs of all chunks in fixed order.
    if (ns == 1) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            int e = tid + s * NT;
            if (e < E) { gacc[2 * e] = acc[s].x; gacc[2 * e + 1] = acc[s].y; }
        }
    } else {
        cplx* part = tile;                               // the tile buffer is free now (>= NT entries)
        __syncthreads();
        part[tid] = (tid / E < ns) ? acc[0] : mk(0.0, 0.0);
        __syncthreads();
        if (tid < E) {
            cplx t = mk(0.0, 0.0);
            for (int sl = 0; sl < ns; sl++) t = cadd(t, part[sl * E + tid]);
            gacc[2 * tid] = t.x; gacc[2 * tid + 1] = t.y;
        }
    }
    __syncthreads();
    double* Gp = G + ((long long)pair * gridDim.x + blockIdx.x) * E * 2;
    for (int i = tid; i < E * 2; i += NT) Gp[i] = gacc[i];
}

// ---------------------------------------------------------------------------------
// (2) Hermitian eigen-solve of each Gram matrix: parallel cyclic two-sided Jacobi.
// grid = npairs.  Writes the accumulated row transformation Q (W_new = Q W_old), the
// diagonal (squared row norms), and re-zeroes G.
//
// One inner round = np disjoint rotations R = diag of 2x2 blocks; G' = R G R^H and
// Q' = R Q are formed in ONE pass from the old matrices into a second shared-memory
// buffer (two barriers per round).  `cross_only`: both 16-row blocks are already
// internally orthogonal (true after the first outer sweep), so only the 16x16 cross
// pairs are rotated: 16 rounds per inner sweep instead of 31.
// ---------------------------------------------------------------------------------
constexpr int GS = PMAX + 1;                              // row stride of the shared matrices
constexpr size_t EIG_SMEM = 2ull * PMAX * GS * sizeof(cplx);   // g, q

constexpr int NTE = 768;   // k_eig block: 256 threads update G, 512 update Q in the same pass
__global__ void __launch_bounds__(NTE)
k_eig(double* __restrict__ G, int nchunks, cplx* __restrict__ Qout, int nrows, double tol2, int max_inner,
      float cross_ratio, int cross_only, PairSpec ps, int slot_base, int single, int* __restrict__ notconv,
      int* __restrict__ rotated, double* __restrict__ sig2, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    extern __shared__ __align__(16) unsigned char eig_smem[];
    cplx* g = (cplx*)eig_smem;                            // [PMAX][GS]
    cplx* q = g + PMAX * GS;                              // [PMAX][GS]
    __shared__ double rcc[PMAX / 2];
    __shared__ cplx roff[PMAX / 2];
    __shared__ int rpp[PMAX / 2], rqq[PMAX / 2], ract[PMAX / 2];
    __shared__ int s_any, s_sweep, s_off, s_mc, s_mi, s_round;
    __shared__ unsigned char sched[(PMAX - 1) * (PMAX / 2) * 2];     // round-robin schedule (p,q) per round/slot
    const int tid = threadIdx.x, pair = slot_base + blockIdx.x;   // workspace slot of this pair
    const int n = nrows, ne = n + (n & 1), np = ne / 2;
    // nchunks > 0: G holds per-chunk partial Gram matrices [pair][chunk][PMAX*PMAX*2] written with plain
    // stores by k_gram_mma (summed here in fixed order); nchunks < 0: single-block path, compact slabs.
    if (nchunks > 0) {
        const double* Gp = G + (long long)pair * nchunks * PMAX * PMAX * 2;
        for (int e = tid; e < n * n; e += NTE) {
            double re = 0.0, im = 0.0;
#pragma unroll 4
            for (int c = 0; c < nchunks; c++) {            // independent L2 loads, fixed summation order
                const double2 v = __ldcg((const double2*)(Gp + (long long)c * PMAX * PMAX * 2 + 2 * e));
                re += v.x; im += v.y;
            }
            int i = e / n, j = e % n;
            g[i * GS + j] = mk(re, im);
            q[i * GS + j] = mk(i == j ? 1.0 : 0.0, 0.0);
        }
    } else {
        // single-block path: -nchunks compact slabs [2 n^2] written by the CTAs of k_gram / k_gram_skinny.
        // Fixed-order sum: value v is summed over slabs sl, sl + nsl, ... by thread (sl, v), then over sl.
        __shared__ double part[NTE];
        const int ns_tot = -nchunks, E2 = 2 * n * n;
        int nsl = NTE / E2;
        if (nsl < 1) nsl = 1;
        if (nsl > ns_tot) nsl = ns_tot;
        const double* Gp = G + (long long)pair * ns_tot * E2;
        double* gd = (double*)g;
        if (nsl == 1) {
            for (int v = tid; v < E2; v += NTE) {
                double sacc = 0.0;
#pragma unroll 8
                for (int c = 0; c < ns_tot; c++) sacc += __ldcg(Gp + (long long)c * E2 + v);
                const int e = v >> 1;
                gd[2 * ((e / n) * GS + (e % n)) + (v & 1)] = sacc;
            }
        } else {
            if (tid < E2 * nsl) {
                const int v = tid % E2, sl = tid / E2;
                double sacc = 0.0;
#pragma unroll 8
                for (int c = sl; c < ns_tot; c += nsl) sacc += __ldcg(Gp + (long long)c * E2 + v);
                part[tid] = sacc;
            }
            __syncthreads();
            for (int v = tid; v < E2; v += NTE) {
                double sacc = 0.0;
                for (int sl = 0; sl < nsl; sl++) sacc += part[sl * E2 + v];
                const int e = v >> 1;
                gd[2 * ((e / n) * GS + (e % n)) + (v & 1)] = sacc;
            }
        }
        for (int e = tid; e < n * n; e += NTE) q[(e / n) * GS + (e % n)] = mk((e / n) == (e % n) ? 1.0 : 0.0, 0.0);
    }
    for (int e = tid; e < (ne - 1) * np; e += NTE) {
        int r = e / np, k = e % np, a, b;
        if (ne == 2) { a = 0; b = 1; } else circle_pair(r, k, ne, a, b);
        sched[2 * e] = (unsigned char)a;
        sched[2 * e + 1] = (unsigned char)b;
    }
    // per-thread work items of the update phase do not depend on the round
    // threads [0,256): one 2x2 block of G each (np*np <= 256); threads [256,768): one (pair, column) item of Q.
    // (updating only the upper blocks k <= l and mirroring them was measured slightly slower)
    const int blk_k = tid / np, blk_l = tid % np;
    const int nblk = np * np;
    const int qt = tid - 256;
    const int q_k0 = qt >= 0 ? qt / n : np, q_c0 = qt >= 0 ? qt % n : 0;
    if (tid == 0) { s_any = 0; s_off = 0; s_mc = 0; s_mi = 0; }
    __syncthreads();
    // Fresh Gram matrix: already diagonal to tolerance?  Largest relative off-diagonal
    // |g_ij|^2/(g_ii g_jj) among cross-block and intra-block entries decides the schedule.
    {
        int offd = 0;
        float mc = 0.f, mi = 0.f;
        for (int e = tid; e < n * n; e += NTE) {
            int i = e / n, j = e % n;
            if (i < j) {
                double a = g[i * GS + i].x, b = g[j * GS + j].x;
                if (a > 0.0 && b > 0.0) {
                    double m2 = cabs2(g[i * GS + j]);
                    if (m2 > tol2 * a * b) {
                        offd = 1;
                        float rel = (float)(m2 / (a * b));
                        if ((i < BSZ) == (j < BSZ)) mi = fmaxf(mi, rel); else mc = fmaxf(mc, rel);
                    }
                }
            }
        }
        if (offd) s_off = 1;
        // non-negative floats order like their bit patterns
        if (mc > 0.f) atomicMax(&s_mc, __float_as_int(mc));
        if (mi > 0.f) atomicMax(&s_mi, __float_as_int(mi));
    }
    __syncthreads();
    // cross-only schedule while the intra-block residual is well below the cross-block one
    const bool cross = cross_only && !single && n == PMAX &&
                       (__int_as_float(s_mi) <= cross_ratio * __int_as_float(s_mc));
    const int nrounds = cross ? BSZ : ne - 1;
    if (s_off) {
        for (int sweep = 0; sweep < max_inner; sweep++) {
            if (tid == 0) s_sweep = 0;
            __syncthreads();
            for (int r = 0; r < nrounds; r++) {
                if (tid == 0) s_round = 0;
                __syncwarp();
                if (tid < np) {
                    int p, qq;
                    if (cross) { p = tid; qq = BSZ + ((tid + r) & (BSZ - 1)); }
                    else { p = sched[2 * (r * np + tid)]; qq = sched[2 * (r * np + tid) + 1]; }
                    bool act = false;
                    double c = 1.0, s = 0.0;
                    cplx u = mk(0.0, 0.0);
                    if (p < n && qq < n) {
                        double a = g[p * GS + p].x, b = g[qq * GS + qq].x;
                        cplx gpq = g[p * GS + qq];
                        double mag2 = cabs2(gpq);
                        if (a > 0.0 && b > 0.0 && mag2 > tol2 * a * b) {
                            // overflow-free form without 1/|g| (a numerically null row shrinks geometrically under
                            // repeated rotations; ((b-a)/2|g|)^2 can then overflow):  dd = (b-a)/2,
                            // den = |dd| + sqrt(dd^2+|g|^2), R = 1/sqrt(den^2+|g|^2):  c = den R,  s u = sign(dd) R g
                            const double dd = 0.5 * (b - a);
                            const double hh = fma(dd, dd, mag2);
                            const double den = fabs(dd) + hh * rsqrt(hh);      // hh > 0 (mag2 > 0); rsqrt is the cheaper chain
                            const double R = rsqrt(fma(den, den, mag2));
                            c = den * R;
                            s = copysign(R, dd);
                            u = gpq;
                            act = true;
                        }
                    }
                    rpp[tid] = p; rqq[tid] = qq; rcc[tid] = c;
                    roff[tid] = mk(-s * u.x, -s * u.y);
                    ract[tid] = act ? 1 : 0;
                    if (act) { s_sweep = 1; s_any = 1; s_round = 1; }
                }
                __syncthreads();
                if (s_round) {
                    // G' = R G R^H by 2x2 blocks: block (k,l) = rows {p_k,q_k} x cols {p_l,q_l} depends only
                    // on the same block of G (4 loads, 4 stores).  In place: every block is owned by one thread.
                    if (tid < nblk) {
                        const int k = blk_k, l = blk_l;
                        const int pk = rpp[k], qk = rqq[k], pl = rpp[l], ql = rqq[l];
                        const bool vk = qk < n, vl = ql < n;           // dummy partner (odd n): single row/col
                        const double ck = rcc[k], cl = rcc[l];
                        const cplx ok = roff[k], ol = roff[l];         // -s u  (row p gets ck*x + ok*y, row q gets -conj(ok)*x + ck*y)
                        cplx g00 = g[pk * GS + pl];
                        cplx g01 = vl ? g[pk * GS + ql] : mk(0.0, 0.0);
                        cplx g10 = vk ? g[qk * GS + pl] : mk(0.0, 0.0);
                        cplx g11 = (vk && vl) ? g[qk * GS + ql] : mk(0.0, 0.0);
                        // rows: [x0;x1] = R_k [g0*; g1*]
                        cplx a00 = cadd(cscale(g00, ck), cmul(ok, g10));
                        cplx a01 = cadd(cscale(g01, ck), cmul(ok, g11));
                        cplx a10 = csub(cscale(g10, ck), cmul(cconj(ok), g00));
                        cplx a11 = csub(cscale(g11, ck), cmul(cconj(ok), g01));
                        // cols: [y0 y1] = [a*0 a*1] R_l^H :  y0 = cl a0 + conj(ol) a1 ; y1 = -ol a0 + cl a1
                        cplx b00 = cadd(cscale(a00, cl), cmulc(a01, ol));
                        cplx b01 = csub(cscale(a01, cl), cmul(ol, a00));
                        cplx b10 = cadd(cscale(a10, cl), cmulc(a11, ol));
                        cplx b11 = csub(cscale(a11, cl), cmul(ol, a10));
                        if (k == l) {
                            b00.y = 0.0; b11.y = 0.0;
                            if (ract[k]) { b01 = mk(0.0, 0.0); b10 = mk(0.0, 0.0); }
                        }
                        g[pk * GS + pl] = b00;
                        if (vl) g[pk * GS + ql] = b01;
                        if (vk) g[qk * GS + pl] = b10;
                        if (vk && vl) g[qk * GS + ql] = b11;
                    }
                    // Q' = R Q : rows p_k, q_k
                    {
                        const int k = q_k0, col = q_c0;
                        if (k < np && ract[k]) {
                        const int pk = rpp[k], qk = rqq[k];
                        const double ck = rcc[k];
                        const cplx ok = roff[k];
                        cplx x = q[pk * GS + col], y = q[qk * GS + col];
                        q[pk * GS + col] = cadd(cscale(x, ck), cmul(ok, y));
                        q[qk * GS + col] = csub(cscale(y, ck), cmul(cconj(ok), x));
                        }
                    }
                }
                __syncthreads();
            }
            if (!s_sweep) break;
            __syncthreads();
        }
    }
    cplx* Qp = Qout + (long long)pair * PMAX * PMAX;
    for (int e = tid; e < n * n; e += NTE) Qp[e] = q[(e / n) * GS + (e % n)];
    if (tid < n) {
        int row = tid;
        if (!single) {
            int bi, bj;
            get_pair(ps, blockIdx.x, bi, bj);
            row = tid < BSZ ? bi * BSZ + tid : bj * BSZ + (tid - BSZ);
        }
        sig2[row] = g[tid * GS + tid].x;
    }
    if (tid == 0) {
        rotated[pair] = s_any;
        if (s_off) {
            atomicAdd(notconv, 1);
            // largest relative off-diagonal^2 met in this sweep (fresh Gram): lets the host skip the
            // verification sweep when the quadratically convergent last sweep started below 1e-9
            int mx = s_mc > s_mi ? s_mc : s_mi;
            atomicMax(notconv + 2, mx);
        }
    }
}

// ---------------------------------------------------------------------------------
// (2') Hermitian eigen-solve of a full 32 x 32 Gram matrix (multi-block mode), round-2 formulation.
// Same arithmetic as k_eig (cyclic two-sided Jacobi, same rotation formula, same schedules, same convergence
// rules); what changes is how a rotation round is laid out on the SM.  Measured on k_eig (ncu, round 1): ~2800
// warp-instructions and ~1.5 us per round at ~1 warp-instruction per cycle -- FP64 issue bound -- with all
// other warps waiting at a barrier while 16 threads compute the round's rotations.  Here
//   * G is kept Hermitian by construction: only the 120 blocks above the diagonal of the 16 x 16 tiling into
//     2 x 2 blocks are rotated (one thread each, 4 warps) and mirrored, plus the 16 diagonal blocks;
//   * a round is ONE barrier: warp 0 rotates the diagonal blocks and the 16 blocks that hold the NEXT round's
//     pivots, then (warp-synchronously) computes the next round's rotations from them, while four warps rotate
//     the off-diagonal blocks and eleven warps apply the current rotations to Q;
//   * G and the rotation table are double-buffered in shared memory (a round reads one copy and writes the
//     other, so the pivot blocks can be rotated both by warp 0 and by their regular owner: identical values).
// ---------------------------------------------------------------------------------
constexpr int NTE3 = 512;
constexpr int NOFF = (PMAX / 2) * (PMAX / 2 - 1) / 2;     // 120 blocks above the diagonal

struct Eig3Smem {
    cplx g[2][PMAX * GS];
    cplx q[PMAX * GS];
    double rc[2][PMAX / 2];
    cplx ro[2][PMAX / 2];        // row p <- c x + o y ;  row q <- -conj(o) x + c y
    int ract[2][PMAX / 2];
    unsigned char sched[(PMAX - 1) * (PMAX / 2) * 2];     // full round-robin schedule: (p, q) per round / slot
    unsigned char slot[(PMAX - 1) * PMAX];                // its inverse: pair slot of row i in round r
    unsigned char blk[NOFF * 2];                          // (k, l), k < l
    int s_off, s_mc, s_mi, s_stop, s_any;
};

// pair (p, q) of slot k in round r
__device__ __forceinline__ void eig3_pair(const Eig3Smem& sm, bool cross, int r, int k, int& p, int& q) {
    if (cross) { p = k; q = BSZ + ((k + r) & (BSZ - 1)); }
    else { p = sm.sched[2 * (r * (PMAX / 2) + k)]; q = sm.sched[2 * (r * (PMAX / 2) + k) + 1]; }
}
// slot of row i in round r
__device__ __forceinline__ int eig3_slot(const Eig3Smem& sm, bool cross, int r, int i) {
    if (cross) return i < BSZ ? i : ((i - BSZ - r) & (BSZ - 1));
    return sm.slot[r * PMAX + i];
}

// G' = R G R^H on block (k, l): rows {pk, qk} x cols {pl, ql}, read from gi, written to g; off-diagonal blocks also
// write the conjugate-transposed block.
__device__ __forceinline__ void eig3_block(const cplx* gi, cplx* g, int pk, int qk, int pl, int ql, double ck, cplx ok,
                                           double cl, cplx ol, bool diag, bool act) {
    const cplx g00 = gi[pk * GS + pl], g01 = gi[pk * GS + ql], g10 = gi[qk * GS + pl], g11 = gi[qk * GS + ql];
    // rows: [x0;x1] = R_k [g0*; g1*]
    const cplx a00 = cadd(cscale(g00, ck), cmul(ok, g10));
    const cplx a01 = cadd(cscale(g01, ck), cmul(ok, g11));
    const cplx a10 = csub(cscale(g10, ck), cmul(cconj(ok), g00));
    const cplx a11 = csub(cscale(g11, ck), cmul(cconj(ok), g01));
    // cols: [y0 y1] = [a*0 a*1] R_l^H :  y0 = cl a0 + conj(ol) a1 ; y1 = -ol a0 + cl a1
    cplx b00 = cadd(cscale(a00, cl), cmulc(a01, ol));
    cplx b01 = csub(cscale(a01, cl), cmul(ol, a00));
    cplx b10 = cadd(cscale(a10, cl), cmulc(a11, ol));
    cplx b11 = csub(cscale(a11, cl), cmul(ol, a10));
    if (diag) {
        b00.y = 0.0; b11.y = 0.0;
        if (act) { b01 = mk(0.0, 0.0); b10 = mk(0.0, 0.0); }
        g[pk * GS + pl] = b00; g[pk * GS + ql] = b01; g[qk * GS + pl] = b10; g[qk * GS + ql] = b11;
    } else {
        g[pk * GS + pl] = b00; g[pk * GS + ql] = b01; g[qk * GS + pl] = b10; g[qk * GS + ql] = b11;
        g[pl * GS + pk] = cconj(b00); g[ql * GS + pk] = cconj(b01); g[pl * GS + qk] = cconj(b10); g[ql * GS + qk] = cconj(b11);
    }
}

// Eigen-solve of ONE pair by the 512 threads of the calling CTA.  Gp: the pair's nchunks partial Gram slabs;
// Qp: the pair's 32 x 32 output; (bi, bj): its row blocks; pair: its workspace slot.
__device__ __forceinline__ void eig3_run(Eig3Smem& sm, const double* Gp, int nchunks, cplx* Qp, double tol2, int max_inner,
                                         float cross_ratio, int cross_only, int bi, int bj, int pair,
                                         int* __restrict__ notconv, int* __restrict__ rotated,
                                         double* __restrict__ sig2) {
    cplx* g = sm.g[0];
    cplx* q = sm.q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int n = PMAX, np = PMAX / 2;
    // ---- load: sum of the per-chunk partial Gram matrices [chunk][PMAX*PMAX*2] in fixed order ----
    {
        constexpr int NE = PMAX * PMAX / NTE3;            // entries per thread
        double re[NE], im[NE];
#pragma unroll
        for (int it = 0; it < NE; it++) { re[it] = 0.0; im[it] = 0.0; }
#pragma unroll 8
        for (int c = 0; c < nchunks; c++) {                // 2 x 8 independent L2 loads in flight per thread
#pragma unroll
            for (int it = 0; it < NE; it++) {
                const double2 v = __ldcg((const double2*)(Gp + (long long)c * PMAX * PMAX * 2 + 2 * (tid + it * NTE3)));
                re[it] += v.x; im[it] += v.y;
            }
        }
#pragma unroll
        for (int it = 0; it < NE; it++) {
            const int e = tid + it * NTE3, i = e / PMAX, j = e % PMAX;
            g[i * GS + j] = mk(re[it], im[it]);
            q[i * GS + j] = mk(i == j ? 1.0 : 0.0, 0.0);
        }
    }
    for (int e = tid; e < (n - 1) * np; e += NTE3) {
        int r = e / np, k = e % np, a, b;
        circle_pair(r, k, n, a, b);
        sm.sched[2 * e] = (unsigned char)a;
        sm.sched[2 * e + 1] = (unsigned char)b;
        sm.slot[r * PMAX + a] = (unsigned char)k;
        sm.slot[r * PMAX + b] = (unsigned char)k;
    }
    if (tid < NOFF) {
        // t -> (k, l), k < l, row-major over the strict upper triangle of the 16 x 16 tiling
        int k = 0, rem = tid;
        while (rem >= np - 1 - k) { rem -= np - 1 - k; k++; }
        sm.blk[2 * tid] = (unsigned char)k;
        sm.blk[2 * tid + 1] = (unsigned char)(k + 1 + rem);
    }
    if (tid == 0) { sm.s_off = 0; sm.s_mc = 0; sm.s_mi = 0; sm.s_stop = 0; sm.s_any = 0; }
    __syncthreads();
    // Fresh Gram matrix: already diagonal to tolerance?  Largest relative off-diagonal
    // |g_ij|^2/(g_ii g_jj) among cross-block and intra-block entries decides the schedule.
    {
        int offd = 0;
        float mc = 0.f, mi = 0.f;
        for (int e = tid; e < n * n; e += NTE3) {
            const int i = e / n, j = e % n;
            if (i < j) {
                const double a = g[i * GS + i].x, b = g[j * GS + j].x;
                if (a > 0.0 && b > 0.0) {
                    const double m2 = cabs2(g[i * GS + j]);
                    if (m2 > tol2 * a * b) {
                        offd = 1;
                        const float rel = (float)(m2 / (a * b));
                        if ((i < BSZ) == (j < BSZ)) mi = fmaxf(mi, rel); else mc = fmaxf(mc, rel);
                    }
                }
            }
        }
        if (offd) sm.s_off = 1;
        // non-negative floats order like their bit patterns
        if (mc > 0.f) atomicMax(&sm.s_mc, __float_as_int(mc));
        if (mi > 0.f) atomicMax(&sm.s_mi, __float_as_int(mi));
    }
    __syncthreads();
    // cross-only schedule while the intra-block residual is well below the cross-block one
    const bool cross = cross_only && (__int_as_float(sm.s_mi) <= cross_ratio * __int_as_float(sm.s_mc));
    const int nrounds = cross ? BSZ : n - 1;
    const int total = sm.s_off ? max_inner * nrounds : 0;

    // rotation of pair slot j for (global) round rr from the current G -> table[rr & 1]
    auto make_rotation = [&](const cplx* gb, int rr, int j) -> int {
        int p, qq;
        eig3_pair(sm, cross, rr % nrounds, j, p, qq);
        const double a = gb[p * GS + p].x, b = gb[qq * GS + qq].x;
        const cplx gpq = gb[p * GS + qq];
        const double mag2 = cabs2(gpq);
        double c = 1.0;
        cplx o = mk(0.0, 0.0);
        int act = 0;
        if (a > 0.0 && b > 0.0 && mag2 > tol2 * a * b) {
            // overflow-free form without 1/|g| (see k_eig)
            const double dd = 0.5 * (b - a);
            const double hh = fma(dd, dd, mag2);
            const double den = fabs(dd) + hh * rsqrt(hh);
            const double R = rsqrt(fma(den, den, mag2));
            const double s = copysign(R, dd);
            c = den * R;
            o = mk(-s * gpq.x, -s * gpq.y);
            act = 1;
        }
        sm.rc[rr & 1][j] = c;
        sm.ro[rr & 1][j] = o;
        sm.ract[rr & 1][j] = act;
        return act;
    };

    int sweep_any = 0;                                     // warp 0: a rotation was active in the current inner sweep
    if (total > 0 && warp == 0) {
        int act = 0;
        if (lane < np) act = make_rotation(g, 0, lane);
        sweep_any = __any_sync(0xffffffffu, act);
    }
    __syncthreads();
    // warp roles: 0 = pivots + next rotations (alone on its scheduler but for Q warps); 1,2,3,5 = off-diagonal blocks;
    // 4, 6..15 = Q
    const int bw = warp == 5 ? 3 : warp - 1;
    const bool is_blk = warp == 1 || warp == 2 || warp == 3 || warp == 5;
    const int qtid = (warp == 4 ? 0 : warp - 5) * 32 + lane;
    int rr = 0;
    for (; rr < total; rr++) {
        const int r = rr % nrounds, cur = rr & 1;
        const cplx* gi = sm.g[cur];
        cplx* go = sm.g[cur ^ 1];
        if (warp == 0) {
            // diagonal blocks (lanes 0-15) and the blocks holding the next round's pivots (lanes 16-31)
            const bool have_next = rr + 1 < total;
            int k, l;
            bool diag = lane < np;
            if (diag) { k = lane; l = lane; }
            else {
                int p2 = 0, q2 = 0;
                eig3_pair(sm, cross, (rr + 1) % nrounds, lane - np, p2, q2);
                const int k1 = eig3_slot(sm, cross, r, p2), k2 = eig3_slot(sm, cross, r, q2);
                k = k1 < k2 ? k1 : k2;
                l = k1 < k2 ? k2 : k1;
            }
            if (diag || (have_next && k != l)) {
                int pk, qk, pl, ql;
                eig3_pair(sm, cross, r, k, pk, qk);
                eig3_pair(sm, cross, r, l, pl, ql);
                eig3_block(gi, go, pk, qk, pl, ql, sm.rc[cur][k], sm.ro[cur][k], sm.rc[cur][l], sm.ro[cur][l], diag,
                           sm.ract[cur][k] != 0);
            }
            __syncwarp();
            if (have_next) {
                if ((rr + 1) % nrounds == 0) {             // an inner sweep just ended
                    if (!sweep_any) { if (lane == 0) sm.s_stop = 1; }
                    sweep_any = 0;
                }
                int act = 0;
                if (lane < np) act = make_rotation(go, rr + 1, lane);
                sweep_any |= __any_sync(0xffffffffu, act);
            }
        } else if (is_blk) {
            // the 120 blocks above the diagonal (the ones warp 0 also rotates get identical values twice)
            const int t = bw * 32 + lane;
            if (t < NOFF) {
                const int k = sm.blk[2 * t], l = sm.blk[2 * t + 1];
                int pk, qk, pl, ql;
                eig3_pair(sm, cross, r, k, pk, qk);
                eig3_pair(sm, cross, r, l, pl, ql);
                eig3_block(gi, go, pk, qk, pl, ql, sm.rc[cur][k], sm.ro[cur][k], sm.rc[cur][l], sm.ro[cur][l], false, false);
            }
        } else {
            // Q' = R Q: 16 pairs x 32 columns over the 352 threads of the eleven Q warps
            for (int item = qtid; item < np * PMAX; item += 352) {
                const int k = item >> 5, col = item & 31;
                if (sm.ract[cur][k]) {
                    int pk, qk;
                    eig3_pair(sm, cross, r, k, pk, qk);
                    const double ck = sm.rc[cur][k];
                    const cplx ok = sm.ro[cur][k];
                    const cplx x = q[pk * GS + col], y = q[qk * GS + col];
                    q[pk * GS + col] = cadd(cscale(x, ck), cmul(ok, y));
                    q[qk * GS + col] = csub(cscale(y, ck), cmul(cconj(ok), x));
                }
            }
        }
        if (tid == 0) {
            int any = 0;
#pragma unroll
            for (int j = 0; j < np; j++) any |= sm.ract[cur][j];
            if (any) sm.s_any = 1;
        }
        __syncthreads();
        if (sm.s_stop) { rr++; break; }
    }
    g = sm.g[rr & 1];                                      // buffer written by the last executed round
    for (int e = tid; e < n * n; e += NTE3) Qp[e] = q[(e / n) * GS + (e % n)];
    if (tid < n) {
        const int row = tid < BSZ ? bi * BSZ + tid : bj * BSZ + (tid - BSZ);
        sig2[row] = g[tid * GS + tid].x;
    }
    if (tid == 0) {
        rotated[pair] = sm.s_any;
        if (sm.s_off) {
            atomicAdd(notconv, 1);
            // largest relative off-diagonal^2 met in this sweep (fresh Gram): lets the host skip the
            // verification sweep when the quadratically convergent last sweep started below 1e-9
            const int mx = sm.s_mc > sm.s_mi ? sm.s_mc : sm.s_mi;
            atomicMax(notconv + 2, mx);
        }
    }
}

// ---------------------------------------------------------------------------------
// (2'') Mixed-precision eigen-solve of a full 32 x 32 Gram matrix (round 2, second step).
//
// ncu on k_eig3 (profiles/ncu_r02_summary.md): a rotation round costs ~1800 cycles and its critical path is one
// warp's chain of ~45 DEPENDENT FP64 instructions (block update -> pivots -> two rsqrt -> rotation) at ~40 cycles
// each; the FP64 pipe itself is 11 % busy.  The chain cannot be made shorter in FP64, but it does not need FP64:
//   * any unitary Q applied to the pair's rows leaves the SVD exact -- the quality of Q only decides how fast the
//     OUTER iteration converges (its Gram matrices are re-formed in FP64 from W every visit);
//   * so the Jacobi recurrence itself (G update -> pivots -> angle) runs in FP32 on a scaled copy of G, with the
//     diagonal tracked in FP64 (differences of nearly equal squared norms decide the angles of close pairs);
//   * each rotation is then made EXACTLY unitary in FP64 from its FP32 tangent t:  c = 1/sqrt(1 + |t|^2),
//     o = -t c (one FP32-seeded third-order rsqrt step), off the critical path, and applied to Q in FP64.
// An FP32 angle is accurate to ~1e-7 of itself, so a visit leaves a residual of max(eps^2, 1e-7 eps) from an
// off-diagonal level eps instead of eps^2: the same number of outer sweeps to reach 1e-14.
// Round layout (one barrier per round, 512 threads):
//   warp 0         diagonal blocks + the 16 blocks holding the next round's pivots, then the next round's angles;
//   warps 1,2,3,5  the 120 off-diagonal blocks (FP32), mirrored;
//   warp 6         lanes 0-15: exact (c, o) of round r in FP64 from t(r);
//   warps 4, 7-15  Q <- R(r-1) Q in FP64 (one round behind, 320 threads for 512 items).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct Eig4Smem {
    cplx gd[PMAX * GS];              // fresh Gram matrix in FP64 (load, convergence scan)
    cplx q[PMAX * GS];
    float2 gf[2][PMAX * GS];         // scaled FP32 working copies (double-buffered)
    double hd[PMAX];                 // HALF the scaled diagonal, FP64
    float rcf[2][PMAX / 2];          // FP32 rotation of a round: c, o
    float2 rof[2][PMAX / 2];
    float2 rt[2][PMAX / 2];          // FP32 tangent t = sign(dd) g / den, kept for the FP64 side (two rounds alive)
    int ract[4][PMAX / 2];           // active flags (G side reads [r&1], Q side one and two rounds later)
    double rc[2][PMAX / 2];          // exact FP64 rotation: row p <- c x + o y ;  row q <- -conj(o) x + c y
    cplx ro[2][PMAX / 2];
    unsigned char sched[(PMAX - 1) * (PMAX / 2) * 2];
    unsigned char slot[(PMAX - 1) * PMAX];
    unsigned char blk[NOFF * 2];
    int s_off, s_mc, s_mi, s_stop, s_any;
    double s_scale;
};

__device__ __forceinline__ void eig4_pair(const Eig4Smem& sm, bool cross, int r, int k, int& p, int& q) {
    if (cross) { p = k; q = BSZ + ((k + r) & (BSZ - 1)); }
    else { p = sm.sched[2 * (r * (PMAX / 2) + k)]; q = sm.sched[2 * (r * (PMAX / 2) + k) + 1]; }
}
__device__ __forceinline__ int eig4_slot(const Eig4Smem& sm, bool cross, int r, int i) {
    if (cross) return i < BSZ ? i : ((i - BSZ - r) & (BSZ - 1));
    return sm.slot[r * PMAX + i];
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 f2mulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)
__device__ __forceinline__ float2 f2conj(float2 a) { return make_float2(a.x, -a.y); }

// FP32 G' = R G R^H on block (k, l), read from gi, written to go (+ mirror for off-diagonal blocks)
__device__ __forceinline__ void eig4_block(const float2* gi, float2* go, int pk, int qk, int pl, int ql, float ck, float2 ok,
                                           float cl, float2 ol, bool diag, bool act) {
    const float2 g00 = gi[pk * GS + pl], g01 = gi[pk * GS + ql], g10 = gi[qk * GS + pl], g11 = gi[qk * GS + ql];
    const float2 u0 = f2mul(ok, g10), u1 = f2mul(ok, g11), u2 = f2mul(f2conj(ok), g00), u3 = f2mul(f2conj(ok), g01);
    const float2 a00 = make_float2(fmaf(g00.x, ck, u0.x), fmaf(g00.y, ck, u0.y));
    const float2 a01 = make_float2(fmaf(g01.x, ck, u1.x), fmaf(g01.y, ck, u1.y));
    const float2 a10 = make_float2(fmaf(g10.x, ck, -u2.x), fmaf(g10.y, ck, -u2.y));
    const float2 a11 = make_float2(fmaf(g11.x, ck, -u3.x), fmaf(g11.y, ck, -u3.y));
    const float2 v0 = f2mulc(a01, ol), v1 = f2mul(ol, a00), v2 = f2mulc(a11, ol), v3 = f2mul(ol, a10);
    float2 b00 = make_float2(fmaf(a00.x, cl, v0.x), fmaf(a00.y, cl, v0.y));
    float2 b01 = make_float2(fmaf(a01.x, cl, -v1.x), fmaf(a01.y, cl, -v1.y));
    float2 b10 = make_float2(fmaf(a10.x, cl, v2.x), fmaf(a10.y, cl, v2.y));
    float2 b11 = make_float2(fmaf(a11.x, cl, -v3.x), fmaf(a11.y, cl, -v3.y));
    if (diag) {
        b00.y = 0.f; b11.y = 0.f;
        if (act) { b01 = make_float2(0.f, 0.f); b10 = b01; }
        go[pk * GS + pl] = b00; go[pk * GS + ql] = b01; go[qk * GS + pl] = b10; go[qk * GS + ql] = b11;
    } else {
        go[pk * GS + pl] = b00; go[pk * GS + ql] = b01; go[qk * GS + pl] = b10; go[qk * GS + ql] = b11;
        go[pl * GS + pk] = f2conj(b00); go[ql * GS + pk] = f2conj(b01); go[pl * GS + qk] = f2conj(b10); go[ql * GS + qk] = f2conj(b11);
    }
}

// Eigen-solve of ONE pair by the 512 threads of the calling CTA (same contract as eig3_run).  Returns false --
// nothing written -- when the squared row norms of the pair span more than 24 decades (numerically null rows of
// a rank-deficient matrix): their mutual cosines would fall below the FP32 range, the caller then runs eig3_run.
__device__ __forceinline__ bool eig4_run(Eig4Smem& sm, const double* Gp, int nchunks, cplx* Qp, double tol2, int max_inner,
                                         float cross_ratio, int cross_only, int bi, int bj, int pair,
                                         int* __restrict__ notconv, int* __restrict__ rotated,
                                         double* __restrict__ sig2, long long* tdbg = nullptr) {
    cplx* g = sm.gd;
    cplx* q = sm.q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int n = PMAX, np = PMAX / 2;
    // ---- load: sum of the per-chunk partial Gram matrices [chunk][PMAX*PMAX*2] in fixed order ----
    {
        constexpr int NE = PMAX * PMAX / NTE3;            // entries per thread
        double re[NE], im[NE];
#pragma unroll
        for (int it = 0; it < NE; it++) { re[it] = 0.0; im[it] = 0.0; }
#pragma unroll 8
        for (int c = 0; c < nchunks; c++) {                // 2 x 8 independent L2 loads in flight per thread
#pragma unroll
            for (int it = 0; it < NE; it++) {
                const double2 v = __ldcg((const double2*)(Gp + (long long)c * PMAX * PMAX * 2 + 2 * (tid + it * NTE3)));
                re[it] += v.x; im[it] += v.y;
            }
        }
#pragma unroll
        for (int it = 0; it < NE; it++) {
            const int e = tid + it * NTE3, i = e / PMAX, j = e % PMAX;
            g[i * GS + j] = mk(re[it], im[it]);
            q[i * GS + j] = mk(i == j ? 1.0 : 0.0, 0.0);
        }
    }
    for (int e = tid; e < (n - 1) * np; e += NTE3) {
        int r = e / np, k = e % np, a, b;
        circle_pair(r, k, n, a, b);
        sm.sched[2 * e] = (unsigned char)a;
        sm.sched[2 * e + 1] = (unsigned char)b;
        sm.slot[r * PMAX + a] = (unsigned char)k;
        sm.slot[r * PMAX + b] = (unsigned char)k;
    }
    if (tid < NOFF) {
        int k = 0, rem = tid;
        while (rem >= np - 1 - k) { rem -= np - 1 - k; k++; }
        sm.blk[2 * tid] = (unsigned char)k;
        sm.blk[2 * tid + 1] = (unsigned char)(k + 1 + rem);
    }
    if (tid == 0) { sm.s_off = 0; sm.s_mc = 0; sm.s_mi = 0; sm.s_stop = 0; sm.s_any = 0; }
    __syncthreads();
    // Fresh Gram matrix (FP64): already diagonal to tolerance?  Largest relative off-diagonal
    // |g_ij|^2/(g_ii g_jj) among cross-block and intra-block entries decides the schedule.
    {
        int offd = 0;
        float mc = 0.f, mi = 0.f;
        for (int e = tid; e < n * n; e += NTE3) {
            const int i = e / n, j = e % n;
            if (i < j) {
                const double a = g[i * GS + i].x, b = g[j * GS + j].x;
                if (a > 0.0 && b > 0.0) {
                    const double m2 = cabs2(g[i * GS + j]);
                    if (m2 > tol2 * a * b) {
                        offd = 1;
                        const float rel = (float)(m2 / (a * b));
                        if ((i < BSZ) == (j < BSZ)) mi = fmaxf(mi, rel); else mc = fmaxf(mc, rel);
                    }
                }
            }
        }
        if (offd) sm.s_off = 1;
        if (mc > 0.f) atomicMax(&sm.s_mc, __float_as_int(mc));
        if (mi > 0.f) atomicMax(&sm.s_mi, __float_as_int(mi));
    }
    // scale: largest diagonal entry -> 1 (FP32 range; every entry of a PSD matrix is bounded by it)
    if (warp == 0) {
        double d = g[lane * GS + lane].x, dn = d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
            dn = fmin(dn, __shfl_xor_sync(0xffffffffu, dn, o));
        }
        if (lane == 0) sm.s_scale = (d > 0.0 && dn > 1e-24 * d) ? 1.0 / d : -1.0;
    }
    __syncthreads();
    if (sm.s_scale < 0.0 && sm.s_off) return false;        // uniform
    if (tdbg && threadIdx.x == 0) tdbg[0] = gtimer();      // loaded + scanned
    const bool cross = cross_only && (__int_as_float(sm.s_mi) <= cross_ratio * __int_as_float(sm.s_mc));
    const int nrounds = cross ? BSZ : n - 1;
    const int total = sm.s_off ? max_inner * nrounds : 0;
    const float tol2f = (float)tol2;
    if (total > 0) {
        const double sc = sm.s_scale;
        for (int e = tid; e < n * n; e += NTE3) {
            const int i = e / n, j = e % n;
            const cplx v = g[i * GS + j];
            sm.gf[0][i * GS + j] = make_float2((float)(v.x * sc), (float)(v.y * sc));
        }
        if (tid < n) sm.hd[tid] = 0.5 * sc * g[tid * GS + tid].x;
    }
    __syncthreads();

    // FP32 angle of pair slot j for (global) round rr from buffer gb and the FP64 half-diagonal; also advances the
    // half-diagonal by the rotation (a' = a - T, b' = b + T, T = Re(conj(t) g)).
    auto make_rotation = [&](const float2* gb, int rr, int j) -> int {
        int p, qq;
        eig4_pair(sm, cross, rr % nrounds, j, p, qq);
        const double ha = sm.hd[p], hb = sm.hd[qq];
        const float2 gpq = gb[p * GS + qq];
        float c = 1.f;
        float2 o = make_float2(0.f, 0.f), t = o;
        int act = 0;
        if (ha > 0.0 && hb > 0.0) {
            // relative test in FP32 on cos = g / sqrt(a b) (a b itself may underflow FP32)
            const float ra = rsqrtf(2.f * (float)ha), rb = rsqrtf(2.f * (float)hb);
            const float cx = gpq.x * ra * rb, cy = gpq.y * ra * rb;
            if (fmaf(cx, cx, cy * cy) > tol2f) {
                float ddf = (float)(hb - ha);                  // dd = (b - a) / 2 from the FP64 half-diagonal
                // local power-of-two scale: the squares below must not underflow for numerically null rows
                const float m = fmaxf(fabsf(ddf), fmaxf(fabsf(gpq.x), fabsf(gpq.y)));
                const int ex = (__float_as_int(m) >> 23) & 0xff;
                const float s2 = __int_as_float((254 - (ex > 0 ? (ex < 253 ? ex : 253) : 1)) << 23);
                ddf *= s2;
                const float gx = gpq.x * s2, gy = gpq.y * s2;
                const float mag2 = fmaf(gx, gx, gy * gy);
                const float hh = fmaf(ddf, ddf, mag2);
                const float den = fabsf(ddf) + sqrtf(hh);
                const float inv = copysignf(__frcp_rn(den), ddf);
                t = make_float2(gx * inv, gy * inv);            // |t| <= 1, scale free
                c = rsqrtf(fmaf(t.x, t.x, fmaf(t.y, t.y, 1.f)));
                o = make_float2(-t.x * c, -t.y * c);
                act = 1;
                // T = Re(conj(t) g) in unscaled units; half-diagonal moves by T / 2
                const double T2 = 0.5 * (double)fmaf(t.x, gpq.x, t.y * gpq.y);
                sm.hd[p] = ha - T2;
                sm.hd[qq] = hb + T2;
            }
        }
        sm.rcf[rr & 1][j] = c;
        sm.rof[rr & 1][j] = o;
        sm.rt[rr & 1][j] = t;
        sm.ract[rr & 3][j] = act;
        return act;
    };

    int sweep_any = 0;
    if (total > 0 && warp == 0) {
        int act = 0;
        if (lane < np) act = make_rotation(sm.gf[0], 0, lane);
        sweep_any = __any_sync(0xffffffffu, act);
    }
    __syncthreads();
    if (tdbg && threadIdx.x == 0) { tdbg[1] = gtimer(); tdbg[4] = total; }      // converted, first angles
    // warp roles: 0 = pivots + next angles; 1,2,3,5 = off-diagonal blocks (FP32); 6 = exact rotations (FP64);
    // 4, 7..15 = Q (FP64), 320 threads
    const int bw = warp == 5 ? 3 : warp - 1;
    const bool is_blk = warp == 1 || warp == 2 || warp == 3 || warp == 5;
    const bool is_q = warp == 4 || warp >= 7;
    const int qtid = (warp == 4 ? 0 : warp - 6) * 32 + lane;
    // iteration `it`: G update of round it (FP32) and the angles of round it+1; exact rotation of round it; Q update
    // of round it-1.  nG = rounds that run (shrinks when an inner sweep ends without a rotation); one drain iteration.
    int nG = total;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};               // timing study (tdbg): cycles per phase, lane 0 of warps 0 / 6 / 7
    for (int it = 0; it < nG + 1; it++) {
        const bool g_live = it < nG;
        const int r = it % nrounds, cur = it & 1;
        const long long ck0 = tdbg ? clock64() : 0;
        if (g_live && warp == 0) {
            const float2* gi = sm.gf[cur];
            float2* go = sm.gf[cur ^ 1];
            const bool have_next = it + 1 < total;
            int k, l;
            const bool diag = lane < np;
            if (diag) { k = lane; l = lane; }
            else {
                int p2 = 0, q2 = 0;
                eig4_pair(sm, cross, (it + 1) % nrounds, lane - np, p2, q2);
                const int k1 = eig4_slot(sm, cross, r, p2), k2 = eig4_slot(sm, cross, r, q2);
                k = k1 < k2 ? k1 : k2;
                l = k1 < k2 ? k2 : k1;
            }
            if (diag || (have_next && k != l)) {
                int pk, qk, pl, ql;
                eig4_pair(sm, cross, r, k, pk, qk);
                eig4_pair(sm, cross, r, l, pl, ql);
                eig4_block(gi, go, pk, qk, pl, ql, sm.rcf[cur][k], sm.rof[cur][k], sm.rcf[cur][l], sm.rof[cur][l], diag,
                           sm.ract[it & 3][k] != 0);
            }
            __syncwarp();
            if (tdbg) cyc[0] += clock64() - ck0;
            if (have_next) {
                if ((it + 1) % nrounds == 0) {             // an inner sweep just ended
                    if (!sweep_any) { if (lane == 0) sm.s_stop = 1; }
                    sweep_any = 0;
                }
                int act = 0;
                if (lane < np) act = make_rotation(go, it + 1, lane);
                sweep_any |= __any_sync(0xffffffffu, act);
            }
            if (tdbg) cyc[1] += clock64() - ck0;
        } else if (g_live && is_blk) {
            const int t = bw * 32 + lane;
            if (t < NOFF) {
                const int k = sm.blk[2 * t], l = sm.blk[2 * t + 1];
                int pk, qk, pl, ql;
                eig4_pair(sm, cross, r, k, pk, qk);
                eig4_pair(sm, cross, r, l, pl, ql);
                eig4_block(sm.gf[cur], sm.gf[cur ^ 1], pk, qk, pl, ql, sm.rcf[cur][k], sm.rof[cur][k], sm.rcf[cur][l],
                           sm.rof[cur][l], false, false);
            }
        } else if (g_live && warp == 6) {
            // exact unitary rotation of round `it` from its FP32 tangent: c = 1/sqrt(1+|t|^2), o = -t c
            if (lane < np) {
                const int j = lane;
                double c = 1.0;
                cplx o = mk(0.0, 0.0);
                if (sm.ract[it & 3][j]) {
                    const float2 t = sm.rt[cur][j];
                    const double tx = (double)t.x, ty = (double)t.y;
                    const double x = fma(tx, tx, fma(ty, ty, 1.0));          // in [1, 2]
                    const double y0 = (double)rsqrtf((float)x);
                    const double e = fma(-x * y0, y0, 1.0);                  // 1 - x y0^2, |e| ~ 2^-22
                    c = fma(y0 * e, fma(0.375, e, 0.5), y0);                 // y0 (1 + e/2 + 3 e^2/8): error ~ e^3
                    o = mk(-tx * c, -ty * c);
                }
                sm.rc[cur][j] = c;
                sm.ro[cur][j] = o;
            }
            if (tdbg) cyc[2] += clock64() - ck0;
        } else if (is_q && it >= 1) {
            // Q' = R(it-1) Q in FP64
            const int rq = it - 1, rr_ = rq % nrounds, cq = rq & 1;
            for (int item = qtid; item < np * PMAX; item += 320) {
                const int k = item >> 5, col = item & 31;
                if (sm.ract[rq & 3][k]) {
                    int pk, qk;
                    eig4_pair(sm, cross, rr_, k, pk, qk);
                    const double ck = sm.rc[cq][k];
                    const cplx ok = sm.ro[cq][k];
                    const cplx x = q[pk * GS + col], y = q[qk * GS + col];
                    q[pk * GS + col] = cadd(cscale(x, ck), cmul(ok, y));
                    q[qk * GS + col] = csub(cscale(y, ck), cmul(cconj(ok), x));
                }
            }
            if (tdbg) cyc[3] += clock64() - ck0;
        }
        if (tid == 32 && g_live) {
            int any = 0;
#pragma unroll
            for (int j = 0; j < np; j++) any |= sm.ract[it & 3][j];
            if (any) sm.s_any = 1;
        }
        __syncthreads();
        if (tdbg) cyc[4] += clock64() - ck0;
        if (g_live && sm.s_stop) nG = it + 1;              // uniform: read after the barrier
    }
    if (tdbg) {
        // lane 0 of warps 0 (block, block+angles, whole iteration), 6 (exact rotation), 7 (Q) report through shared memory
        double* scr = (double*)sm.gd;                      // the FP64 Gram copy is dead by now (sig2 reads it only if !s_any)
        __syncthreads();
        if (tid == 0) { scr[600] = (double)cyc[0]; scr[601] = (double)cyc[1]; scr[604] = (double)cyc[4]; }
        if (tid == 6 * 32) scr[602] = (double)cyc[2];
        if (tid == 7 * 32) scr[603] = (double)cyc[3];
        __syncthreads();
        if (tid == 0) { tdbg[6] = (long long)scr[600]; tdbg[7] = (long long)scr[601]; tdbg[8] = (long long)scr[602];
                        tdbg[9] = (long long)scr[603]; tdbg[10] = (long long)scr[604]; }
    }
    if (tdbg && threadIdx.x == 0) { tdbg[2] = gtimer(); tdbg[5] = nG; }         // rotations done
    for (int e = tid; e < n * n; e += NTE3) Qp[e] = q[(e / n) * GS + (e % n)];
    if (tid < n) {
        const int row = tid < BSZ ? bi * BSZ + tid : bj * BSZ + (tid - BSZ);
        // rotated rows: the FP64-tracked diagonal (approximate, not final); untouched pairs: the fresh Gram diagonal
        sig2[row] = (sm.s_any && sm.s_scale > 0.0) ? 2.0 * sm.hd[tid] / sm.s_scale : g[tid * GS + tid].x;
    }
    if (tid == 0) {
        rotated[pair] = sm.s_any;
        if (sm.s_off) {
            atomicAdd(notconv, 1);
            const int mx = sm.s_mc > sm.s_mi ? sm.s_mc : sm.s_mi;
            atomicMax(notconv + 2, mx);
        }
    }
    return true;
}

__global__ void __launch_bounds__(NTE3)
k_eig3(double* __restrict__ G, int nchunks, cplx* __restrict__ Qout, double tol2, int max_inner, float cross_ratio,
       int cross_only, PairSpec ps, int slot_base, int* __restrict__ notconv, int* __restrict__ rotated,
       double* __restrict__ sig2, const int* __restrict__ done, int mixed) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    extern __shared__ __align__(16) unsigned char eig_smem[];
    const int pair = slot_base + blockIdx.x;
    int bi, bj;
    get_pair(ps, blockIdx.x, bi, bj);
    if (mixed && eig4_run(*reinterpret_cast<Eig4Smem*>(eig_smem), G + (long long)pair * nchunks * PMAX * PMAX * 2, nchunks,
                          Qout + (long long)pair * PMAX * PMAX, tol2, max_inner, cross_ratio, cross_only, bi, bj, pair,
                          notconv, rotated, sig2))
        return;
    __syncthreads();
    {
        eig3_run(*reinterpret_cast<Eig3Smem*>(eig_smem), G + (long long)pair * nchunks * PMAX * PMAX * 2, nchunks,
                 Qout + (long long)pair * PMAX * PMAX, tol2, max_inner, cross_ratio, cross_only, bi, bj, pair, notconv,
                 rotated, sig2);
    }
}

// ---------------------------------------------------------------------------------
// DMMA versions of (1) and (3) for full 32-row pairs (multi-block mode).  Fragments are
// loaded straight from global memory: a lane's 16-byte load of W[row][col] carries the
// re and im parts that serve as A and B operands of mma.sync.m8n8k4.f64.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// G = W_pair W_pair^H over a column chunk.  Each of the 8 warps owns two of the sixteen 8x8 output
// tiles (tile row w/2, tile columns 2(w%2), 2(w%2)+1) and runs over the whole chunk, so there is no
// cross-warp reduction: the warp stores its tiles straight into the chunk's partial slab, which
// k_eig sums over chunks.  A lane's 16-byte load W[row][k] is both an A and a B operand.
// One 8-warp team: partial Gram of the pair's 32 rows over columns [c0, c1) -> slab Gp [PMAX*PMAX*2] (plain stores).
template <bool CG>
__device__ __forceinline__ cplx ldw_(const cplx* p) { return CG ? __ldcg(p) : *p; }

// One 8 x 8 tile (mt, nt) of W_pair W_pair^H accumulated over the columns [k0, k1) (operands past k1 are zero).
template <bool CG>
__device__ __forceinline__ void gram_tile_acc(const cplx* pa, const cplx* pb, long long k0, long long k1, int t,
                                              double (&cr)[2], double (&ci)[2]) {
    // register double-buffering, prefetch distance PF k-steps (operands come from L2)
    constexpr int PF = 4;
    cplx fa[PF], fb[PF];
#pragma unroll
    for (int s = 0; s < PF; s++) {
        const long long col = k0 + 4 * s + t;
        const bool ok = col < k1;
        fa[s] = ok ? ldw_<CG>(pa + col) : mk(0.0, 0.0);
        fb[s] = ok ? ldw_<CG>(pb + col) : mk(0.0, 0.0);
    }
    for (long long k = k0; k < k1; k += 4 * PF) {
#pragma unroll
        for (int s = 0; s < PF; s++) {
            const cplx wa = fa[s], wb = fb[s];
            const long long col = k + 4 * (PF + s) + t;           // same slot, PF steps ahead
            const bool ok = col < k1;
            fa[s] = ok ? ldw_<CG>(pa + col) : mk(0.0, 0.0);
            fb[s] = ok ? ldw_<CG>(pb + col) : mk(0.0, 0.0);
            // W_m conj(W_n): re = ar*br + ai*bi ; im = ai*br - ar*bi   (zero operands past k1 add nothing)
            dmma884(cr[0], cr[1], wa.x, wb.x);
            dmma884(ci[0], ci[1], wa.y, wb.x);
            dmma884(cr[0], cr[1], wa.y, wb.y);
            dmma884(ci[0], ci[1], -wa.x, wb.y);
        }
    }
}

// Partial Gram matrix of a pair over the columns [c0, c1) by ONE 8-warp team, Hermitian symmetry used: only the 10
// upper 8 x 8 tiles are computed (37.5 % fewer DMMAs than the 16 tiles of round 1).  Balance: the 10 tiles x 4 column
// quarters = 40 units go to the warps five at a time in tile-major order, so a warp accumulates the tail quarters
// [s, 4) of tile A = 5w/4 and the head quarters [0, s] of tile A + 1 (s = 5w mod 4).  A tile shared by two warps is
// finished by the warp that holds its tail: the other one deposits its partial sum in shared memory (`dep`, 8 x 64
// complex per team; fixed order own + deposit).  The finished tile and its conjugate mirror are stored into the slab,
// so the eigen-solvers read a full 32 x 32 matrix as before.  Contains one __syncthreads(): every thread of the CTA
// must call it.
template <bool CG>
__device__ __forceinline__ void gram_mma_part(const cplx* W, long long ldw, int bi, int bj, long long c0,
                                              long long c1, double* __restrict__ Gp, int warp, int lane, cplx* dep) {
    const int g = lane >> 2, t = lane & 3;
    auto rowptr = [&](int r) { return W + (long long)(r < BSZ ? bi * BSZ + r : bj * BSZ + (r - BSZ)) * ldw; };
    auto tile_mn = [](int ti, int& m, int& n) {           // upper tiles, row-major: (0,0) (0,1) (0,2) (0,3) (1,1) ...
        m = ti < 4 ? 0 : (ti < 7 ? 1 : (ti < 9 ? 2 : 3));
        n = ti < 4 ? ti : (ti < 7 ? ti - 3 : (ti < 9 ? ti - 5 : 3));
    };
    long long qlen = ((c1 - c0 + 3) / 4 + 3) / 4 * 4;      // quarter length, a multiple of the k-step
    if (qlen < 4) qlen = 4;
    auto qcol = [&](int qd) { const long long c = c0 + qd * qlen; return c < c1 ? c : c1; };
    const int sq = (5 * warp) & 3, ta = (5 * warp) >> 2, tb = ta + 1;
    int ma, na, mb, nb;
    tile_mn(ta, ma, na);
    tile_mn(tb < 10 ? tb : 9, mb, nb);
    double ar[2] = {0.0, 0.0}, ai[2] = {0.0, 0.0}, br[2] = {0.0, 0.0}, bim[2] = {0.0, 0.0};
    gram_tile_acc<CG>(rowptr(ma * 8 + g), rowptr(na * 8 + g), qcol(sq), qcol(4), t, ar, ai);
    gram_tile_acc<CG>(rowptr(mb * 8 + g), rowptr(nb * 8 + g), qcol(0), qcol(sq + 1), t, br, bim);
    auto store_tile = [&](int m, int n, const double* xr, const double* xi) {
        const int row = m * 8 + g, col = n * 8 + 2 * t;
        *(double4*)(Gp + ((long long)row * PMAX + col) * 2) = make_double4(xr[0], xi[0], xr[1], xi[1]);
        if (m != n) {                                      // mirror: G[col][row] = conj(G[row][col])
            *(double2*)(Gp + ((long long)col * PMAX + row) * 2) = make_double2(xr[0], -xi[0]);
            *(double2*)(Gp + ((long long)(col + 1) * PMAX + row) * 2) = make_double2(xr[1], -xi[1]);
        }
    };
    if (sq == 3) store_tile(mb, nb, br, bim);              // the head quarters were all four: tile B is complete
    else {
        dep[warp * 64 + 2 * lane] = mk(br[0], bim[0]);
        dep[warp * 64 + 2 * lane + 1] = mk(br[1], bim[1]);
    }
    __syncthreads();
    if (sq != 0) {                                         // tile A's head quarters came from the previous warp
        const cplx d0 = dep[(warp - 1) * 64 + 2 * lane], d1 = dep[(warp - 1) * 64 + 2 * lane + 1];
        ar[0] += d0.x; ai[0] += d0.y; ar[1] += d1.x; ai[1] += d1.y;
    }
    store_tile(ma, na, ar, ai);
}

// G = W_pair W_pair^H over a column chunk.  Each of the 8 warps owns two of the sixteen 8x8 output
// tiles (tile row w/2, tile columns 2(w%2), 2(w%2)+1) and runs over the whole chunk, so there is no
// cross-warp reduction: the warp stores its tiles straight into the chunk's partial slab, which
// k_eig sums over chunks.  A lane's 16-byte load W[row][k] is both an A and a B operand.
__global__ void __launch_bounds__(NT)
k_gram_mma(const cplx* __restrict__ W, long long ldw, int len, int chunk, PairSpec ps, int slot_base,
           double* __restrict__ G, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    const int tid = threadIdx.x, pair = slot_base + blockIdx.y;
    int bi, bj;
    get_pair(ps, blockIdx.y, bi, bj);
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < len) ? c0 + chunk : len;
    __shared__ cplx dep[8 * 64];
    gram_mma_part<false>(W, ldw, bi, bj, c0, c1, G + ((long long)pair * gridDim.x + blockIdx.x) * PMAX * PMAX * 2,
                         tid >> 5, tid & 31, dep);
}

// One 8-warp team: rows <- Q rows over columns [c0, c1): each warp owns 8-column strips (stride 64), loads the 32x8
// strip as B fragments, multiplies by Q (A fragments from the padded shared planes qr / qi) and stores in place.
constexpr int QS = PMAX + 4;
template <bool CG>
__device__ __forceinline__ void apply_mma_part(cplx* W, long long ldw, int bi, int bj, long long c0,
                                               long long c1, const double* qr, const double* qi, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // software pipeline: the next strip's B fragments are in flight while the current one is multiplied
    long long n0 = c0 + warp * 8;
    cplx b[8];
    {
        long long col = n0 + g;
        bool ok = (n0 < c1) && (col < c1);
#pragma unroll
        for (int kt = 0; kt < 8; kt++) {
            int r = kt * 4 + t;
            long long row = r < BSZ ? bi * BSZ + r : bj * BSZ + (r - BSZ);
            b[kt] = ok ? ldw_<CG>(W + row * ldw + col) : mk(0.0, 0.0);
        }
    }
    while (n0 < c1) {
        const long long n1 = n0 + 64;
        cplx bn[8];
        {
            long long col = n1 + g;
            bool ok = (n1 < c1) && (col < c1);
#pragma unroll
            for (int kt = 0; kt < 8; kt++) {
                int r = kt * 4 + t;
                long long row = r < BSZ ? bi * BSZ + r : bj * BSZ + (r - BSZ);
                bn[kt] = ok ? ldw_<CG>(W + row * ldw + col) : mk(0.0, 0.0);
            }
        }
        double cr[4][2], ci[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) { cr[mt][0] = cr[mt][1] = ci[mt][0] = ci[mt][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < 8; kt++) {
            // eight independent accumulators back to back, then their second products (same order per accumulator)
            double ar[4], ai[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                ar[mt] = qr[(mt * 8 + g) * QS + kt * 4 + t];
                ai[mt] = qi[(mt * 8 + g) * QS + kt * 4 + t];
            }
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                dmma884(cr[mt][0], cr[mt][1], ar[mt], b[kt].x);
                dmma884(ci[mt][0], ci[mt][1], ar[mt], b[kt].y);
            }
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                dmma884(cr[mt][0], cr[mt][1], -ai[mt], b[kt].y);
                dmma884(ci[mt][0], ci[mt][1], ai[mt], b[kt].x);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            int r = mt * 8 + g;
            long long row = r < BSZ ? bi * BSZ + r : bj * BSZ + (r - BSZ);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                long long c2 = n0 + 2 * t + e;
                if (c2 < c1) W[row * ldw + c2] = mk(cr[mt][e], ci[mt][e]);
            }
        }
#pragma unroll
        for (int kt = 0; kt < 8; kt++) b[kt] = bn[kt];
        n0 = n1;
    }
}

// Wext[rows] <- Q Wext[rows]
__global__ void __launch_bounds__(NT, 2)
k_apply_mma(cplx* __restrict__ W, long long ldw, long long lenx, int chunk, PairSpec ps, int slot_base,
            const cplx* __restrict__ Q, const int* __restrict__ rotated, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    const int pair = slot_base + blockIdx.y;
    if (!rotated[pair]) return;
    __shared__ double qr[PMAX * QS], qi[PMAX * QS];
    const int tid = threadIdx.x;
    int bi, bj;
    get_pair(ps, blockIdx.y, bi, bj);
    const cplx* Qp = Q + (long long)pair * PMAX * PMAX;
    for (int e = tid; e < PMAX * PMAX; e += NT) {
        cplx v = Qp[e];
        qr[(e / PMAX) * QS + (e % PMAX)] = v.x;
        qi[(e / PMAX) * QS + (e % PMAX)] = v.y;
    }
    __syncthreads();
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < lenx) ? c0 + chunk : lenx;
    apply_mma_part<false>(W, ldw, bi, bj, c0, c1, qr, qi, tid >> 5, tid & 31);
}

// ---------------------------------------------------------------------------------
// Fused Jacobi round (round 2).  The three kernels of a round (Gram, eigen-solve, update) and the two
// cross-stream hand-overs between them cost more than the work itself on matrices of the MPS path (a 1024 x 1024
// round: ~7 us of Gram, ~25 us of eigen-solve, ~10 us of update, but ~60 us per round measured; 6 launches and 8
// event calls per round on the host).  Here a round is ONE cooperative launch over all pairs:
//   grid = (chunks, pairs), 512 threads = two 8-warp teams.
//   1. each team forms the pair's partial Gram matrix over its sub-chunk of columns -> slab (plain stores);
//   2. ticket per pair: the LAST CTA of a pair to arrive solves the 32 x 32 eigen-problem (eig3_run), writes Q,
//      and releases the pair's flag (epoch = round number); the other CTAs of the pair spin on the flag
//      (cooperative launch: every CTA is resident, so the wait cannot deadlock);
//   3. each team applies Q to its sub-chunk of columns.
// Reproducible: the slabs are summed in fixed order whichever CTA arrives last.
// ---------------------------------------------------------------------------------
constexpr int NTR = 512;
struct RoundSmem {
    union {
        Eig3Smem e;
        Eig4Smem e4;
        struct { double qr[PMAX * QS], qi[PMAX * QS]; } a;
        cplx dep[2 * 8 * 64];                                 // Gram phase: tile partials passed between warps
    };
    int s_last;
};
constexpr size_t EIG34_SMEM = sizeof(Eig3Smem) > sizeof(Eig4Smem) ? sizeof(Eig3Smem) : sizeof(Eig4Smem);

// grid-wide barrier of a cooperative launch (all CTAs resident): monotonically increasing counter, no reset
__device__ __forceinline__ void round_grid_barrier(unsigned int* bar, unsigned int& target, unsigned int nblocks) {
    __syncthreads();
    target += nblocks;
    if (threadIdx.x == 0) {
        __threadfence();                                   // publish this CTA's rows of W
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

// Rounds [r_begin, r_end) of one sweep in ONE cooperative launch (a grid barrier between rounds: the launch gap of
// ~10 us per round was a fifth of the round).  W is read with ld.global.cg throughout: other SMs rewrite it between
// rounds of the same launch.
template <bool DBG>
__global__ void __launch_bounds__(NTR, 1)
k_round(cplx* W, long long ldw, int len, long long lenx, int chunk, int nbp, int r_begin, int r_end,
        double* G, cplx* Q, double tol2, int max_inner, float cross_ratio, int cross_only, int* notconv, int* rotated,
        double* sig2, unsigned int* ticket, unsigned int* flag, unsigned int* bar, unsigned int* ver,
        unsigned int epoch0, unsigned int bar0, int mixed, long long* dbg) {
    // dbg (QM_ROUND_DEBUG=1): per-CTA sums of phase durations in ns: [0] gram, [1] ticket, [2] wait (waiters), [11] eig
    // (solvers), [3] update, [4..8] inside the eigen-solve (load+scan, convert, rotations, write-back, rounds),
    // [9] rounds, [10] eigen-solves, [12..15] cycles of the mixed-precision loop phases
    extern __shared__ __align__(16) unsigned char eig_smem[];
    RoundSmem& sm = *reinterpret_cast<RoundSmem*>(eig_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, team = warp >> 3, w8 = warp & 7;
    const int pair = blockIdx.y, nch = gridDim.x, nslab = 2 * nch;
    const unsigned int nblocks = gridDim.x * gridDim.y;
    double* Gp = G + (long long)pair * nslab * PMAX * PMAX * 2;
    cplx* Qp = Q + (long long)pair * PMAX * PMAX;
    unsigned int bar_target = bar0;
    for (int r = r_begin; r < r_end; r++) {
        long long tt[4] = {0, 0, 0, 0}, te[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (DBG && tid == 0) tt[0] = gtimer();
        const PairSpec ps = {0, r, nbp, 0, 0};
        const unsigned int epoch = epoch0 + (unsigned int)(r - r_begin);
        int bi, bj;
        get_pair(ps, pair, bi, bj);
        // Dataflow between rounds (ver != NULL): this CTA owns the columns [2 x chunk x blockIdx.x, + 2 chunk) of its
        // pair's rows in both phases, so round e may start as soon as the two CTAs that updated the same columns of
        // row blocks bi and bj in round e-1 are done -- ver[chunk index][block] = last completed epoch -- instead of
        // waiting for the slowest pair of the whole grid (ncu: 23 % of the warp samples at the grid barrier).
        if (ver && epoch > 1u) {
            if (tid < 2) {
                const unsigned int* vp = ver + (long long)blockIdx.x * nbp + (tid == 0 ? bi : bj);
                unsigned int v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(vp) : "memory");
                } while ((int)(v - (epoch - 1u)) < 0);
            }
            __syncthreads();
        }
        // ---- 1. partial Gram matrices (teams whose columns lie in the identity extension write a zero slab) ----
        const int sub = 2 * blockIdx.x + team;
        const long long cc0 = (long long)sub * chunk;
        const long long cc1 = (cc0 + chunk < lenx) ? cc0 + chunk : lenx;
        {
            const long long c0 = cc0 < len ? cc0 : len;
            const long long c1 = cc1 < len ? cc1 : len;
            gram_mma_part<true>(W, ldw, bi, bj, c0, c1 > c0 ? c1 : c0, Gp + (long long)sub * PMAX * PMAX * 2, w8, lane,
                                sm.dep + team * 8 * 64);
        }
        __syncthreads();
        if (DBG && tid == 0) tt[1] = gtimer();
        if (tid == 0) {
            __threadfence();                               // the CTA's slabs before its ticket
            const unsigned int t = atomicAdd(&ticket[pair], 1u);
            sm.s_last = (t == (unsigned int)nch - 1u);
            if (sm.s_last) ticket[pair] = 0u;              // every CTA of the pair has drawn its ticket
        }
        __syncthreads();
        if (DBG && tid == 0) tt[2] = gtimer();
        const bool was_last = sm.s_last != 0;
        if (was_last) {
            // ---- 2. eigen-solve by the last CTA of the pair ----
            __threadfence();
            if (!(mixed && eig4_run(sm.e4, Gp, nslab, Qp, tol2, max_inner, cross_ratio, cross_only, bi, bj, pair, notconv,
                                    rotated, sig2, DBG ? te : nullptr))) {
                __syncthreads();
                eig3_run(sm.e, Gp, nslab, Qp, tol2, max_inner, cross_ratio, cross_only, bi, bj, pair, notconv, rotated, sig2);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag + pair), "r"(epoch) : "memory");
        } else {
            if (tid == 0) {
                unsigned int v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag + pair) : "memory");
                } while (v != epoch);
            }
        }
        __syncthreads();
        if (DBG && tid == 0) tt[3] = gtimer();
        // ---- 3. update ----
        if (__ldcg(rotated + pair)) {
            for (int e = tid; e < PMAX * PMAX; e += NTR) {
                const cplx v = __ldcg(Qp + e);
                sm.a.qr[(e / PMAX) * QS + (e % PMAX)] = v.x;
                sm.a.qi[(e / PMAX) * QS + (e % PMAX)] = v.y;
            }
            __syncthreads();
            apply_mma_part<true>(W, ldw, bi, bj, cc0, cc1 > cc0 ? cc1 : cc0, sm.a.qr, sm.a.qi, w8, lane);
        }
        if (DBG) {
            __syncthreads();
            if (tid == 0) {
                const long long t4 = gtimer();
                long long* d = dbg + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
                d[0] += tt[1] - tt[0]; d[1] += tt[2] - tt[1]; d[was_last ? 11 : 2] += tt[3] - tt[2]; d[3] += t4 - tt[3]; d[9] += 1;
                if (was_last && te[0]) {
                    d[4] += te[0] - tt[2]; d[5] += te[1] - te[0]; d[6] += te[2] - te[1]; d[7] += tt[3] - te[2];
                    d[8] += te[5]; d[10] += 1; d[12] += te[6]; d[13] += te[7]; d[14] += te[8]; d[15] += te[9];
                }
            }
        }
        if (ver) {
            __syncthreads();                               // both teams done with their columns (and the shared-memory union)
            if (tid == 0) {
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ver + (long long)blockIdx.x * nbp + bi), "r"(epoch) : "memory");
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ver + (long long)blockIdx.x * nbp + bj), "r"(epoch) : "memory");
            }
        } else if (r + 1 < r_end) {
            round_grid_barrier(bar, bar_target, nblocks);  // rows of W and the shared-memory union
        }
    }
}

// ---------------------------------------------------------------------------------
// (3) Row update  Wext[rows] <- Q * Wext[rows].  grid = (nchunks, npairs).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
k_apply(cplx* __restrict__ W, long long ldw, long long lenx, int chunk, int round, int nbp, int single,
        int nrows, const cplx* __restrict__ Q, const int* __restrict__ rotated, const int* __restrict__ done) {
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
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
// conj bit 1: conjugate; bit 2: scale by 1/S[j]^2 instead of 1/S[j]
__global__ void k_transpose(cplx* __restrict__ out, long long ldo, const cplx* __restrict__ in, long long ldi,
                            const int* __restrict__ perm, const double* __restrict__ S, int conj, int nsel,
                            long long len, int na) {
    const int ssq = conj & 2;
    conj &= 1;
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
            if (S) {
                double s = S[j];
                if (ssq) s *= s;
                v = cscale(v, (s > 0.0) ? 1.0 / s : 0.0);
            }
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

// Narrow variants of the transpose for the skinny TT-SVD splits (2^i x 2r with 2r <= 32): the
// 32x32 tile kernel would use 2r of 32 tile columns.  Here one thread owns one long-axis index, so
// both sides are coalesced: the narrow side is contiguous per thread (and per warp when the leading
// dimension equals the width), the long side is contiguous across the warp.
// out[a][j] = f(in[perm[j]][a]),  j < nsel (long, one thread each),  a < len <= 32
__global__ void k_transpose_narrow_in(cplx* __restrict__ out, long long ldo, const cplx* __restrict__ in,
                                      long long ldi, const int* __restrict__ perm, const double* __restrict__ S,
                                      int conj, long long nsel, int len) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nsel;
         j += (long long)gridDim.x * blockDim.x) {
        const long long src = perm ? perm[j] : j;
        double sc = 1.0;
        if (S) { double sv = S[j]; sc = (sv > 0.0) ? 1.0 / sv : 0.0; }
        const cplx* p = in + src * ldi;
        for (int a = 0; a < len; a++) {
            cplx v = p[a];
            if (conj) v.y = -v.y;
            out[(long long)a * ldo + j] = cscale(v, sc);
        }
    }
}
// out[a][j] = f(in[perm[j]][a]),  j < nsel <= 32,  a < len (long, one thread each)
__global__ void k_transpose_narrow_out(cplx* __restrict__ out, long long ldo, const cplx* __restrict__ in,
                                       long long ldi, const int* __restrict__ perm, const double* __restrict__ S,
                                       int conj, int nsel, long long len) {
    __shared__ long long srow[32];
    __shared__ double ssc[32];
    if (threadIdx.x < nsel) {
        srow[threadIdx.x] = (long long)(perm ? perm[threadIdx.x] : threadIdx.x) * ldi;
        double sc = 1.0;
        if (S) { double sv = S[threadIdx.x]; sc = (sv > 0.0) ? 1.0 / sv : 0.0; }
        ssc[threadIdx.x] = sc;
    }
    __syncthreads();
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < len;
         a += (long long)gridDim.x * blockDim.x) {
        cplx* q = out + a * ldo;
        for (int j = 0; j < nsel; j++) {
            cplx v = in[srow[j] + a];
            if (conj) v.y = -v.y;
            q[j] = cscale(v, ssc[j]);
        }
    }
}

// Skinny Gram / update for <= 4 short vectors (first TT-SVD splits: 2 or 4 rows of 2^22..2^23
// amplitudes): one column per thread, everything in registers, one pass over W at HBM speed.
__global__ void __launch_bounds__(NT)
k_gram_skinny(const cplx* __restrict__ W, long long ldw, long long len, int nrows, double* __restrict__ G,
              const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double red[NT / 32][32];
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < len;
         c += (long long)gridDim.x * blockDim.x) {
        cplx w[4];
#pragma unroll
        for (int r = 0; r < 4; r++) w[r] = (r < nrows) ? W[(long long)r * ldw + c] : mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                acc[2 * (i * 4 + j)] += w[i].x * w[j].x + w[i].y * w[j].y;        // w_i conj(w_j)
                acc[2 * (i * 4 + j) + 1] += w[i].y * w[j].x - w[i].x * w[j].y;
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = warp_sum(acc[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 32; i++) red[warp][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        double ssum = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NT / 32; w2++) ssum += red[w2][threadIdx.x];
        int e = threadIdx.x >> 1, i = e >> 2, j = e & 3;
        // compact slab of this CTA (summed in fixed order by k_eig)
        if (i < nrows && j < nrows) G[(long long)blockIdx.x * 2 * nrows * nrows + 2 * (i * nrows + j) + (threadIdx.x & 1)] = ssum;
    }
}

__global__ void __launch_bounds__(NT)
k_apply_skinny(cplx* __restrict__ W, long long ldw, long long lenx, int nrows, const cplx* __restrict__ Q,
               const int* __restrict__ rotated, const int* __restrict__ done) {
    if (done && *done) return;
    if (!rotated[0]) return;
    __shared__ cplx qs[16];
    if (threadIdx.x < nrows * nrows) qs[(threadIdx.x / nrows) * 4 + (threadIdx.x % nrows)] = Q[threadIdx.x];
    __syncthreads();
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < lenx;
         c += (long long)gridDim.x * blockDim.x) {
        cplx w[4];
#pragma unroll
        for (int r = 0; r < 4; r++) w[r] = (r < nrows) ? W[(long long)r * ldw + c] : mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < nrows) {
                cplx a = mk(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j < nrows) cfma(a, qs[i * 4 + j], w[j]);
                W[(long long)i * ldw + c] = a;
            }
        }
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

// static mode bookkeeping: end of a sweep / end of the SVD
__global__ void k_sweep_end(int* __restrict__ nc, float early2) {
    // converged: nothing above tol in this sweep, or everything already below the early threshold
    // (the rotations just applied take it to ~early2, i.e. far below tol)
    if (nc[0] == 0 || __int_as_float(nc[2]) <= early2) nc[1] = 1;
    nc[0] = 0;
    nc[2] = 0;
}
__global__ void k_static_check(const int* __restrict__ nc, int* __restrict__ mismatch) {
    if (!nc[1]) mismatch[0] = 1;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int MAXCH = 32;     // Gram partial slots per pair (DMMA path)
constexpr int MAXCH1 = 64;    // Gram slabs of the single-block path

struct Work {
    cplx* W; double* G; cplx* Q; int* rotated; double* sig2; int* perm; int* notconv; cplx* T; unsigned int* sync;
    size_t total;
};

Work carve(const Geom& g, void* base) {
    Work w;
    size_t off = 0;
    char* b = (char*)base;
    w.W = (cplx*)(b + off); off += align_up((size_t)g.nvp * g.ldw * sizeof(cplx));
    w.T = nullptr;
    if (!g.ext) { w.T = (cplx*)(b + off); off += align_up((size_t)g.len * g.nv * sizeof(cplx)); }   // back-multiplication operand
    {   // multi-block: MAXCH slabs per pair; single block: up to MAXCH1 full slabs or 148*4 skinny ones
        size_t slabs = (size_t)g.npairs * MAXCH;
        if (slabs < (size_t)MAXCH1) slabs = MAXCH1;
        w.G = (double*)(b + off); off += align_up(slabs * PMAX * PMAX * 2 * sizeof(double));
    }
    w.Q = (cplx*)(b + off); off += align_up((size_t)g.npairs * PMAX * PMAX * sizeof(cplx));
    w.rotated = (int*)(b + off); off += align_up((size_t)g.npairs * sizeof(int));
    w.sig2 = (double*)(b + off); off += align_up((size_t)g.nvp * sizeof(double));
    w.perm = (int*)(b + off); off += align_up((size_t)g.nvp * sizeof(int));
    w.notconv = (int*)(b + off); off += align_up(4 * sizeof(int));   // [0] not-converged count, [1] done flag, [2] max rel off-diag^2 (float bits)
    // fused rounds: ticket[npairs], flag[npairs], grid barrier, ver[MAXCH/2][nbp] (dataflow between rounds)
    w.sync = (unsigned int*)(b + off);
    off += align_up((2 * (size_t)g.npairs + 1 + (size_t)(MAXCH / 2) * (g.nbp > 0 ? g.nbp : 1)) * sizeof(unsigned int));
    w.total = off;
    return w;
}

}  // namespace

extern "C" long long qm_svd_work_bytes(int m, int n) {
    Geom g0 = make_geom(m, n, 0), g1 = make_geom(m, n, 1);
    Work w0 = carve(g0, nullptr), w1 = carve(g1, nullptr);
    return (long long)(w0.total > w1.total ? w0.total : w1.total);
}

// A (m x n, row-major, lda) is not modified.  U: m x k (ldu), S: k, Vh: k x n (ldvh), k = min(m,n).
// U or Vh may be NULL.  info_host (optional, host int[2]) receives {sweeps, converged}.
static int svd_impl(int m, int n, const void* A_, long long lda, void* U_, long long ldu, void* S_, void* Vh_,
                    long long ldvh, void* work, long long work_bytes, double tol, int max_sweeps,
                    int* info_host, int fixed_sweeps, int* mismatch, int flags, void* stream_) {
    if (m <= 0 || n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    static const int force_acc = getenv("QM_SVD_ACC") ? atoi(getenv("QM_SVD_ACC")) : 0;   // A/B: 1 = always accumulate
    Geom g = make_geom(m, n, ((flags & QM_SVD_BACKMULT) && !force_acc) ? 1 : 0);
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
        if (n <= 32) {
            int nb = ceil_div(m, 256) > 148 * 16 ? 148 * 16 : ceil_div(m, 256);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose_narrow_in<<<nb, 256, 0, st>>>(
                w.W, g.ldw, A, lda, nullptr, nullptr, 0, (long long)m, n));
        } else {
            int na = ceil_div(n, 32);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)na * ceil_div(m, 32)), dim3(32, 8), 0, st>>>(
                w.W, g.ldw, A, lda, nullptr, nullptr, 0, m, n, na));
        }
    }
    QM_CHECK_LAUNCH();
    {
        dim3 grid(ceil_div(g.nvp > 256 ? g.nvp : 256, 256), g.nvp);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_init_ext<<<grid, 256, 0, st>>>(w.W, g.ldw, g.nv, g.ext ? g.nvp : 0, g.len));
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
    // DMMA kernels: aim at ~3 CTAs per SM in flight for a launch of `npl` pairs; chunks in units of one
    // CTA step (32 / 64 columns)
    auto pick_chunk_mma = [&](long long cols, int unit, int npl) {
        long long want = (3LL * 148 + npl - 1) / npl;
        long long chunk = (cols + want - 1) / want;
        chunk = (chunk + unit - 1) / unit * unit;
        if (chunk < unit) chunk = unit;
        return chunk;
    };
    auto gram_chunk = [&](int npl) {
        long long c = pick_chunk_mma(g.len, 32, npl);
        if (c < 64) c = 64;
        if ((g.len + c - 1) / c > MAXCH) c = ((g.len + MAXCH - 1) / MAXCH + 31) / 32 * 32;
        return c;
    };
    const double tol2 = tol * tol;
    static bool eig_attr_set = false;
    static int eig_version = 3;        // 32-row pairs: 3 = FP64 (eig3_run, default), 4 = mixed precision (eig4_run: measured slower in situ), 1 = round-1 kernel
    if (!eig_attr_set) {
        QM_CUDA(cudaFuncSetAttribute(k_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EIG_SMEM));
        QM_CUDA(cudaFuncSetAttribute(k_eig3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EIG34_SMEM));
        if (getenv("QM_EIG")) eig_version = atoi(getenv("QM_EIG"));
        eig_attr_set = true;
    }
    // schedule knobs (defaults chosen from the sweep study in profiles/; env overrides for experiments)
    static int tune_inner0 = -1, tune_inner = 1, tune_cross = 1, tune_groups = 1;
    static float tune_ratio = 1.0f;   // scan in scripts/svd_grid.sh (1e-2 .. 1e4): 10 is fastest on random matrices but stalls on graded spectra
    if (tune_inner0 < 0) {
        const char* e3 = getenv("QM_SVD_CROSS_RATIO");
        if (e3) tune_ratio = (float)atof(e3);
        const char* e0 = getenv("QM_SVD_INNER0");
        const char* e1 = getenv("QM_SVD_INNER");
        const char* e2 = getenv("QM_SVD_CROSS");
        const char* e4 = getenv("QM_SVD_GROUPS");
        tune_inner0 = e0 ? atoi(e0) : 1;      // (2 in round 1: with the fused rounds one inner sweep is 4 % faster, same outer sweeps)
        tune_inner = e1 ? atoi(e1) : 1;
        tune_cross = e2 ? atoi(e2) : 1;
        tune_groups = e4 ? atoi(e4) : 1;
    }
    int sweeps = 0, converged = 0;
    const bool is_static = fixed_sweeps > 0;
    const int* donep = is_static ? w.notconv + 1 : nullptr;
    if (is_static) max_sweeps = fixed_sweeps;
    QM_CUDA(cudaMemsetAsync(w.notconv, 0, 4 * sizeof(int), st));
    QM_CUDA(cudaMemsetAsync(w.sync, 0, (2 * (size_t)g.npairs + 1 + (size_t)(MAXCH / 2) * (g.nbp > 0 ? g.nbp : 1)) * sizeof(unsigned int), st));
    unsigned int fused_barriers = 0;                       // grid barriers passed so far in this SVD
    // a sweep that STARTS below `early` (largest |cos| between rows) ends near early^2 (quadratic regime): no
    // verification sweep after it.  QM_SVD_EARLY overrides for experiments.
    static const float early = getenv("QM_SVD_EARLY") ? (float)atof(getenv("QM_SVD_EARLY")) : 1e-9f;
    const float early2 = early * early;

    // Grouped schedule (multi-block, eager mode): the eigen-solve of a round is latency bound on npairs SMs
    // (~35 us) while the Gram / update kernels fill the GPU.  The block set is split so that two independent
    // pair streams A and B exist (PairSpec); all DMMA kernels run in order on the caller's stream as
    //   gram_A(r), apply_B(r-1), gram_B(r), apply_A(r), gram_A(r+1), ...
    // and the eigen-solves run on side streams, each hidden behind the other groups' DMMA work (2 or 4 groups).  Not used
    // under the event profiler (kernel classes are timed in isolation there) nor in the static / graph mode.
    constexpr int MAXG = 4;
    static cudaStream_t side[MAXG] = {nullptr, nullptr, nullptr, nullptr};
    static cudaEvent_t evG[MAXG], evE[MAXG];
    // 4 streams need nbp % 8 == 0 (even quarter groups that split into halves), 2 streams nbp % 4 == 0
    int NG = 0;
    if (tune_groups && !g.single && !is_static && !qm_prof_active()) {
        if (tune_groups >= 4 && g.nbp >= 64 && (g.nbp % 8) == 0) NG = 4;
        else if (g.nbp >= 32 && (g.nbp % 4) == 0) NG = 2;
    }
    const bool grouped = NG > 0;
    if (grouped && !side[0]) {
        for (int i = 0; i < MAXG; i++) {
            QM_CUDA(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
            QM_CUDA(cudaEventCreateWithFlags(&evG[i], cudaEventDisableTiming));
            QM_CUDA(cudaEventCreateWithFlags(&evE[i], cudaEventDisableTiming));
        }
    }
    struct Phase { int mode, n, rounds, np, offa[MAXG], offb[MAXG]; };
    Phase phases[7];
    int nphases = 0;
    if (NG == 2) {
        const int h = g.nbp / 2, q = h / 2;
        phases[0] = {0, h, h - 1, h / 2, {0, h, 0, 0}, {0, 0, 0, 0}};
        phases[1] = {1, q, q, q, {0, q, 0, 0}, {h, h + q, 0, 0}};
        phases[2] = {1, q, q, q, {0, q, 0, 0}, {h + q, h, 0, 0}};
        nphases = 3;
    } else if (NG == 4) {
        // four quarter groups: their own tournaments, then the three pairings of the quarters, each pairing
        // as two rounds of half x half cross products
        const int s4 = g.nbp / 4, q = s4 / 2;
        phases[nphases++] = {0, s4, s4 - 1, s4 / 2, {0, s4, 2 * s4, 3 * s4}, {0, 0, 0, 0}};
        const int pairing[3][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2}};   // (a1,b1), (a2,b2)
        for (int t = 0; t < 3; t++) {
            const int a1 = pairing[t][0] * s4, b1 = pairing[t][1] * s4, a2 = pairing[t][2] * s4, b2 = pairing[t][3] * s4;
            phases[nphases++] = {1, q, q, q, {a1, a1 + q, a2, a2 + q}, {b1, b1 + q, b2, b2 + q}};
            phases[nphases++] = {1, q, q, q, {a1, a1 + q, a2, a2 + q}, {b1 + q, b1, b2 + q, b2}};
        }
    }
    // Fused schedule (eager mode, multi-block): one cooperative launch per round over all pairs (k_round).
    static int sched_fused = -1, round_capacity = 0;
    if (sched_fused < 0) {
        const char* e5 = getenv("QM_SVD_SCHED");
        sched_fused = !(e5 && strcmp(e5, "grouped") == 0);
        int dev = 0, n_sm = 0, per_sm = 0;
        QM_CUDA(cudaGetDevice(&dev));
        QM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        QM_CUDA(cudaFuncSetAttribute(k_round<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RoundSmem)));
        QM_CUDA(cudaFuncSetAttribute(k_round<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RoundSmem)));
        QM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_round<false>, NTR, sizeof(RoundSmem)));
        round_capacity = n_sm * per_sm;
    }
    int fused_nch = 0;
    if (sched_fused && !g.single && !is_static && g.npairs <= round_capacity) {
        fused_nch = round_capacity / g.npairs;
        if (fused_nch > MAXCH / 2) fused_nch = MAXCH / 2;               // 2 slabs per CTA, MAXCH slabs per pair
        const int by_len = g.len / 128 > 1 ? g.len / 128 : 1;           // at least 64 Gram columns per 8-warp team
        if (fused_nch > by_len) fused_nch = by_len;
    }
    const bool fused = fused_nch > 0;
    // QM_ROUND_DEBUG=1: per-CTA phase timings of k_round accumulated over this SVD, printed at its end
    static const int round_debug = getenv("QM_ROUND_DEBUG") ? atoi(getenv("QM_ROUND_DEBUG")) : 0;
    static long long* round_dbg_buf = nullptr;
    long long* round_dbg = nullptr;
    if (round_debug && fused) {
        if (!round_dbg_buf) QM_CUDA(cudaMalloc(&round_dbg_buf, 4096 * 16 * sizeof(long long)));
        round_dbg = round_dbg_buf;
        QM_CUDA(cudaMemsetAsync(round_dbg, 0, 4096 * 16 * sizeof(long long), st));
    }
    const int npl = grouped ? g.npairs / NG : g.npairs;         // pairs per DMMA launch
    long long chunk_g = g.single ? pick_chunk(g.len) : gram_chunk(npl);
    if (g.single && (g.len + chunk_g - 1) / chunk_g > MAXCH1) chunk_g = ((g.len + MAXCH1 - 1) / MAXCH1 + TC - 1) / TC * TC;
    const long long chunk_a = g.single ? pick_chunk(lenx) : pick_chunk_mma(lenx, 64, npl);
    const int ncg = ceil_div(g.len, chunk_g), nca = ceil_div(lenx, chunk_a);
    const bool skinny = g.single && g.nrows <= 4 && g.len >= 4096;
    const int nb_skinny = ceil_div(g.len, NT) > 148 * 4 ? 148 * 4 : ceil_div(g.len, NT);
    const int eig_chunks = g.single ? -(skinny ? nb_skinny : ncg) : ncg;      // < 0: compact slabs of the single-block path

    for (; sweeps < max_sweeps;) {
        const int max_inner = (sweeps == 0) ? tune_inner0 : tune_inner;
        const int cross_only = (sweeps > 0 && tune_cross) ? 1 : 0;
        auto launch_gram = [&](const PairSpec& ps, int np, int slot, cudaStream_t s) {
            if (is_static) QM_LAUNCH(QM_CLS_SVD_GRAM, s, k_gram_mma<<<dim3(ncg, np), NT, 0, s>>>(
                w.W, g.ldw, g.len, (int)chunk_g, ps, slot, w.G, donep));
            else QM_LAUNCH(QM_CLS_SVD_GRAM, s, qm_launch_dep(k_gram_mma, dim3(ncg, np), dim3(NT), 0, s,
                w.W, g.ldw, g.len, (int)chunk_g, ps, slot, w.G, donep));
        };
        auto launch_eig = [&](const PairSpec& ps, int np, int slot, cudaStream_t s) {
            if (eig_version >= 3 && !g.single) {
                const int mixed = eig_version == 4;
                if (is_static) QM_LAUNCH(QM_CLS_SVD_EIG, s, k_eig3<<<np, NTE3, EIG34_SMEM, s>>>(
                    w.G, eig_chunks, w.Q, tol2, max_inner, tune_ratio, cross_only, ps, slot, w.notconv, w.rotated,
                    w.sig2, donep, mixed));
                else QM_LAUNCH(QM_CLS_SVD_EIG, s, qm_launch_dep(k_eig3, dim3(np), dim3(NTE3), EIG34_SMEM, s,
                    w.G, eig_chunks, w.Q, tol2, max_inner, tune_ratio, cross_only, ps, slot, w.notconv, w.rotated,
                    w.sig2, donep, mixed));
                return;
            }
            if (is_static) QM_LAUNCH(QM_CLS_SVD_EIG, s, k_eig<<<np, NTE, EIG_SMEM, s>>>(
                w.G, eig_chunks, w.Q, g.nrows, tol2, g.single ? 12 : max_inner, tune_ratio, cross_only, ps, slot,
                g.single, w.notconv, w.rotated, w.sig2, donep));
            else QM_LAUNCH(QM_CLS_SVD_EIG, s, qm_launch_dep(k_eig, dim3(np), dim3(NTE), EIG_SMEM, s,
                w.G, eig_chunks, w.Q, g.nrows, tol2, g.single ? 12 : max_inner, tune_ratio, cross_only, ps, slot,
                g.single, w.notconv, w.rotated, w.sig2, donep));
        };
        auto launch_apply = [&](const PairSpec& ps, int np, int slot, cudaStream_t s) {
            if (is_static) QM_LAUNCH(QM_CLS_SVD_APPLY, s, k_apply_mma<<<dim3(nca, np), NT, 0, s>>>(
                w.W, g.ldw, lenx, (int)chunk_a, ps, slot, w.Q, w.rotated, donep));
            else QM_LAUNCH(QM_CLS_SVD_APPLY, s, qm_launch_dep(k_apply_mma, dim3(nca, np), dim3(NT), 0, s,
                w.W, g.ldw, lenx, (int)chunk_a, ps, slot, w.Q, w.rotated, donep));
        };
        if (fused) {
            const int nteams = 2 * fused_nch;
            int chunk = (int)(((lenx + nteams - 1) / nteams + 63) / 64 * 64);       // same columns in Gram and update
            static const int dataflow = !(getenv("QM_SVD_DATAFLOW") && atoi(getenv("QM_SVD_DATAFLOW")) == 0);
            unsigned int* ticket = w.sync;
            unsigned int* flag = w.sync + g.npairs;
            {
                // every round of the sweep in one cooperative launch (QM_SVD_ROUNDS_PER_LAUNCH limits it for A/B runs)
                static const int rpl = getenv("QM_SVD_ROUNDS_PER_LAUNCH") ? atoi(getenv("QM_SVD_ROUNDS_PER_LAUNCH")) : 1 << 20;
                unsigned int* bar = w.sync + 2 * g.npairs;
                unsigned int* ver = dataflow ? w.sync + 2 * g.npairs + 1 : nullptr;
                const unsigned int nblocks = (unsigned int)fused_nch * (unsigned int)g.npairs;
                for (int r0 = 0; r0 < g.rounds; r0 += rpl) {
                    int r1 = r0 + rpl < g.rounds ? r0 + rpl : g.rounds;
                    int rb = r0;
                    unsigned int epoch0 = (unsigned int)(sweeps * g.rounds + r0 + 1);
                    unsigned int bar0 = fused_barriers * nblocks;
                    fused_barriers += (unsigned int)(r1 - r0 - 1);
                    cplx* Wp = w.W; long long ldw = g.ldw; int len = g.len; long long lx = lenx; int nbp = g.nbp;
                    double* Gp = w.G; cplx* Qp = w.Q; double t2 = tol2; int mi = max_inner; float cr = tune_ratio;
                    int co = cross_only; int* nc = w.notconv; int* rot = w.rotated; double* s2 = w.sig2;
                    int mixed = eig_version == 4;
                    long long* dbg = round_dbg;
                    void* args[] = {&Wp, &ldw, &len, &lx, &chunk, &nbp, &rb, &r1, &Gp, &Qp, &t2, &mi, &cr, &co, &nc, &rot,
                                    &s2, &ticket, &flag, &bar, &ver, &epoch0, &bar0, &mixed, &dbg};
                    QM_LAUNCH(QM_CLS_SVD_ROUND, st, cudaLaunchCooperativeKernel(
                        dbg ? (void*)k_round<true> : (void*)k_round<false>, dim3(fused_nch, g.npairs), dim3(NTR), args,
                        sizeof(RoundSmem), st));
                }
            }
            // algorithmic flops: Hermitian Gram = the 10 upper tiles of 16 (a ZHERK), full 32 x 32 update
            qm_prof_work(QM_CLS_SVD_ROUND,
                         8.0 * g.nrows * g.nrows * (0.625 * (double)g.len + (double)lenx) * g.npairs * g.rounds);
        } else if (grouped) {
            for (int ph = 0; ph < nphases; ph++) {
                const Phase& P = phases[ph];
                bool pending[MAXG] = {false, false, false, false};
                PairSpec pspec[MAXG];
                for (int r = 0; r < P.rounds; r++) {
                    for (int grp = 0; grp < NG; grp++) {
                        const PairSpec ps = {P.mode, r, P.n, P.offa[grp], P.offb[grp]};
                        const int slot = grp * P.np;
                        launch_gram(ps, P.np, slot, st);
                        QM_CUDA(cudaEventRecord(evG[grp], st));
                        QM_CUDA(cudaStreamWaitEvent(side[grp], evG[grp], 0));
                        launch_eig(ps, P.np, slot, side[grp]);
                        QM_CUDA(cudaEventRecord(evE[grp], side[grp]));
                        const int o = (grp + 1) % NG;           // the oldest pending update fills the wait
                        if (pending[o]) {
                            QM_CUDA(cudaStreamWaitEvent(st, evE[o], 0));
                            launch_apply(pspec[o], P.np, o * P.np, st);
                            pending[o] = false;
                        }
                        pending[grp] = true;
                        pspec[grp] = ps;
                    }
                }
                for (int k = 0; k < NG; k++) {
                    const int grp = (k + 1) % NG;               // same order as inside the loop
                    if (pending[grp]) {
                        QM_CUDA(cudaStreamWaitEvent(st, evE[grp], 0));
                        launch_apply(pspec[grp], P.np, grp * P.np, st);
                    }
                }
            }
        } else {
            for (int r = 0; r < g.rounds; r++) {
                const PairSpec ps = {0, r, g.nbp, 0, 0};
                if (skinny) {
                    QM_LAUNCH(QM_CLS_SVD_GRAM, st, k_gram_skinny<<<nb_skinny, NT, 0, st>>>(w.W, g.ldw, (long long)g.len, g.nrows,
                                                                                         w.G, donep));
                } else if (g.single) {
                    QM_LAUNCH(QM_CLS_SVD_GRAM, st, k_gram<<<dim3(ncg, g.npairs), NT, 0, st>>>(
                        w.W, g.ldw, g.len, (int)chunk_g, r, g.nbp, g.single, g.nrows, w.G, donep));
                } else {
                    launch_gram(ps, g.npairs, 0, st);
                }
                launch_eig(ps, g.npairs, 0, st);
                if (g.single && g.nrows <= 4 && g.len >= 4096) {
                    int nb = ceil_div(lenx, NT) > 148 * 8 ? 148 * 8 : ceil_div(lenx, NT);
                    QM_LAUNCH(QM_CLS_SVD_APPLY, st, k_apply_skinny<<<nb, NT, 0, st>>>(w.W, g.ldw, lenx, g.nrows, w.Q,
                                                                                    w.rotated, donep));
                } else if (g.single) {
                    QM_LAUNCH(QM_CLS_SVD_APPLY, st, k_apply<<<dim3(nca, g.npairs), NT, 0, st>>>(
                        w.W, g.ldw, lenx, (int)chunk_a, r, g.nbp, g.single, g.nrows, w.Q, w.rotated, donep));
                } else {
                    launch_apply(ps, g.npairs, 0, st);
                }
            }
        }
        // complex MAC = 8 flops: Gram nrows^2 x len, update nrows^2 x lenx, per pair and round
        if (!fused) {
            qm_prof_work(QM_CLS_SVD_GRAM, 0.625 * 8.0 * g.nrows * g.nrows * (double)g.len * g.npairs * g.rounds);
            qm_prof_work(QM_CLS_SVD_APPLY, 8.0 * g.nrows * g.nrows * (double)lenx * g.npairs * g.rounds);
        }
        QM_CHECK_LAUNCH();
        sweeps++;
        if (is_static) {
            // no host round trip: the remaining sweeps become empty launches once `done` is set
            QM_LAUNCH(QM_CLS_SMALL, st, k_sweep_end<<<1, 1, 0, st>>>(w.notconv, early2));
            continue;
        }
        int h[4] = {0, 0, 0, 0};
        QM_CUDA(cudaMemcpyAsync(h, w.notconv, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        QM_CUDA(cudaStreamSynchronize(st));
        QM_CUDA(cudaMemsetAsync(w.notconv, 0, 4 * sizeof(int), st));
        float mx;
        memcpy(&mx, &h[2], sizeof(float));
        if (h[0] == 0 || mx <= early2) { converged = 1; break; }
    }
    if (round_dbg) {
        const int ncta = fused_nch * g.npairs;
        long long* h = (long long*)malloc((size_t)ncta * 16 * sizeof(long long));
        QM_CUDA(cudaMemcpyAsync(h, round_dbg, (size_t)ncta * 16 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        QM_CUDA(cudaStreamSynchronize(st));
        double sum[16] = {0};
        for (int c = 0; c < ncta; c++)
            for (int k = 0; k < 16; k++) sum[k] += (double)h[c * 16 + k];
        free(h);
        const double nl = sum[9] > 0 ? sum[9] : 1, ne = sum[10] > 0 ? sum[10] : 1, nw = nl - ne > 0 ? nl - ne : 1;
        fprintf(stderr, "[k_round %dx%d nch=%d sweeps=%d] per CTA-launch (us): gram %.2f ticket %.2f wait %.2f (waiters) eig-total %.2f (solvers) update %.2f | "
                        "eig: load+scan %.2f convert %.2f rotations %.2f write+flag %.2f rounds %.1f | cycles/round: w0 block %.0f w0 block+angles %.0f prep %.0f Q %.0f\n",
                m, n, fused_nch, sweeps, sum[0] / nl * 1e-3, sum[1] / nl * 1e-3, sum[2] / nw * 1e-3, sum[11] / ne * 1e-3, sum[3] / nl * 1e-3,
                sum[4] / ne * 1e-3, sum[5] / ne * 1e-3, sum[6] / ne * 1e-3, sum[7] / ne * 1e-3, sum[8] / ne,
                sum[12] / (sum[8] > 0 ? sum[8] : 1), sum[13] / (sum[8] > 0 ? sum[8] : 1), sum[14] / (sum[8] > 0 ? sum[8] : 1),
                sum[15] / (sum[8] > 0 ? sum[8] : 1));
    }
    if (is_static && mismatch) QM_LAUNCH(QM_CLS_SMALL, st, k_static_check<<<1, 1, 0, st>>>(w.notconv, mismatch));
    if (info_host) { info_host[0] = sweeps; info_host[1] = converged; }

    // --- sort, emit ---
    QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_sort<<<1, 1024, (size_t)g.nv * sizeof(double), st>>>(w.sig2, g.nv, S, w.perm));
    QM_CHECK_LAUNCH();
    const int k = g.nv;
    cplx* U = (cplx*)U_;
    cplx* Vh = (cplx*)Vh_;
    if (m < n) {
        // U[a][j] = conj(J[perm[j]][a]);  Vh[j][c] = W[perm[j]][c] / S[j]
        if (U && g.ext)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(m, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                U, ldu, w.W + g.len, g.ldw, w.perm, nullptr, 1, k, m, ceil_div(m, 32)));
        if (U && !g.ext) {
            // rows of W are sigma_j z_j:  U = A Z^H Sigma^-1 = A T,  T[a][j] = conj(W[perm[j]][a]) / S[j]^2  (n x k)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(n, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                w.T, k, w.W, g.ldw, w.perm, S, 3, k, n, ceil_div(n, 32)));
            QM_CHECK_LAUNCH();
            int e = qm_zgemm(m, k, n, 1.0, 0.0, A, lda, w.T, k, 0.0, 0.0, U, ldu, 1, 0, 0, 0, 0, stream_);
            if (e) return e;
        }
        if (Vh) {
            dim3 grid(ceil_div(n, 256) > 4096 ? 4096 : ceil_div(n, 256), k);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_rowcopy<<<grid, 256, 0, st>>>(Vh, ldvh, w.W, g.ldw, w.perm, S, 0, k, n));
        }
    } else {
        // U[a][j] = W[perm[j]][a] / S[j];  Vh[j][c] = conj(J[perm[j]][c])
        if (U && k <= 32) {
            int nb = ceil_div(m, 256) > 148 * 16 ? 148 * 16 : ceil_div(m, 256);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose_narrow_out<<<nb, 256, 0, st>>>(
                U, ldu, w.W, g.ldw, w.perm, S, 0, k, (long long)m));
        } else if (U)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(m, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                U, ldu, w.W, g.ldw, w.perm, S, 0, k, m, ceil_div(m, 32)));
        if (Vh && g.ext) {
            dim3 grid(ceil_div(n, 256), k);
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_rowcopy<<<grid, 256, 0, st>>>(Vh, ldvh, w.W + g.len, g.ldw, w.perm, nullptr, 1, k, n));
        }
        if (Vh && !g.ext) {
            // rows of W are sigma_j u_j^T:  Vh = Sigma^-1 U^H A = T^H A,  T[a][j] = W[perm[j]][a] / S[j]^2  (m x k)
            QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)ceil_div(m, 32) * ceil_div(k, 32)), dim3(32, 8), 0, st>>>(
                w.T, k, w.W, g.ldw, w.perm, S, 2, k, m, ceil_div(m, 32)));
            QM_CHECK_LAUNCH();
            int e = qm_zgemm(k, n, m, 1.0, 0.0, w.T, k, A, lda, 0.0, 0.0, Vh, ldvh, 1, 0, 0, 0, 1, stream_);
            if (e) return e;
        }
    }
    QM_CHECK_LAUNCH();
    return 0;
}

extern "C" int qm_svd(int m, int n, const void* A, long long lda, void* U, long long ldu, void* S, void* Vh,
                      long long ldvh, void* work, long long work_bytes, double tol, int max_sweeps,
                      int* info_host, int flags, void* stream) {
    return svd_impl(m, n, A, lda, U, ldu, S, Vh, ldvh, work, work_bytes, tol, max_sweeps, info_host, 0, nullptr,
                    flags, stream);
}

// Sync-free variant (CUDA-graph capturable): exactly `fixed_sweeps` sweeps are enqueued, kernels of the
// sweeps after convergence return immediately, and mismatch[0] is set to 1 if the SVD had not
// converged by then (the caller re-runs that problem through qm_svd).
extern "C" int qm_svd_static(int m, int n, const void* A, long long lda, void* U, long long ldu, void* S, void* Vh,
                             long long ldvh, void* work, long long work_bytes, double tol, int fixed_sweeps,
                             void* mismatch, int flags, void* stream) {
    return svd_impl(m, n, A, lda, U, ldu, S, Vh, ldvh, work, work_bytes, tol, fixed_sweeps, nullptr,
                    fixed_sweeps, (int*)mismatch, flags, stream);
}

// out (cols x rows, ldo) = transpose of in (rows x cols, ldi), optionally conjugated.  The layout
// kernels of the TT-SVD, exposed for tests and bandwidth measurements.
extern "C" int qm_transpose(void* out, long long ldo, const void* in, long long ldi, long long rows, long long cols,
                            int conj, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (rows <= 0 || cols <= 0) return 0;
    if (cols <= 32) {
        int nb = ceil_div(rows, 256) > 148 * 16 ? 148 * 16 : ceil_div(rows, 256);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose_narrow_in<<<nb, 256, 0, st>>>(
            (cplx*)out, ldo, (const cplx*)in, ldi, nullptr, nullptr, conj, rows, (int)cols));
    } else if (rows <= 32) {
        int nb = ceil_div(cols, 256) > 148 * 16 ? 148 * 16 : ceil_div(cols, 256);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose_narrow_out<<<nb, 256, 0, st>>>(
            (cplx*)out, ldo, (const cplx*)in, ldi, nullptr, nullptr, conj, (int)rows, cols));
    } else {
        int na = ceil_div(cols, 32);
        QM_LAUNCH(QM_CLS_SVD_LAYOUT, st, k_transpose<<<(unsigned)((long long)na * ceil_div(rows, 32)), dim3(32, 8), 0, st>>>(
            (cplx*)out, ldo, (const cplx*)in, ldi, nullptr, nullptr, conj, (int)rows, cols, na));
    }
    qm_prof_work(QM_CLS_SVD_LAYOUT, 32.0 * (double)rows * (double)cols);
    QM_CHECK_LAUNCH();
    return 0;
}
