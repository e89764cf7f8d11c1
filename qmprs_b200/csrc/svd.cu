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

__global__ void __launch_bounds__(NTE3)
k_eig3(double* __restrict__ G, int nchunks, cplx* __restrict__ Qout, double tol2, int max_inner, float cross_ratio,
       int cross_only, PairSpec ps, int slot_base, int* __restrict__ notconv, int* __restrict__ rotated,
       double* __restrict__ sig2, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    extern __shared__ __align__(16) unsigned char eig_smem[];
    Eig3Smem& sm = *reinterpret_cast<Eig3Smem*>(eig_smem);
    cplx* g = sm.g[0];
    cplx* q = sm.q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pair = slot_base + blockIdx.x;
    constexpr int n = PMAX, np = PMAX / 2;
    // ---- load: sum of the per-chunk partial Gram matrices [pair][chunk][PMAX*PMAX*2] in fixed order ----
    {
        const double* Gp = G + (long long)pair * nchunks * PMAX * PMAX * 2;
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
    cplx* Qp = Qout + (long long)pair * PMAX * PMAX;
    for (int e = tid; e < n * n; e += NTE3) Qp[e] = q[(e / n) * GS + (e % n)];
    if (tid < n) {
        int bi, bj;
        get_pair(ps, blockIdx.x, bi, bj);
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
__global__ void __launch_bounds__(NT)
k_gram_mma(const cplx* __restrict__ W, long long ldw, int len, int chunk, PairSpec ps, int slot_base,
           double* __restrict__ G, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pair = slot_base + blockIdx.y;
    const int g = lane >> 2, t = lane & 3;
    int bi, bj;
    get_pair(ps, blockIdx.y, bi, bj);
    const int mt = warp >> 1, nt0 = (warp & 1) * 2;
    auto rowptr = [&](int r) { return W + (long long)(r < BSZ ? bi * BSZ + r : bj * BSZ + (r - BSZ)) * ldw; };
    const cplx* pa = rowptr(mt * 8 + g);
    const cplx* pb0 = rowptr(nt0 * 8 + g);
    const cplx* pb1 = rowptr(nt0 * 8 + 8 + g);
    double cr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, ci[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < len) ? c0 + chunk : len;
    // register double-buffering, prefetch distance PF k-steps (operands come from L2)
    constexpr int PF = 4;
    cplx fa[PF], fb0[PF], fb1[PF];
#pragma unroll
    for (int s = 0; s < PF; s++) {
        const long long col = c0 + 4 * s + t;
        const bool ok = col < c1;
        fa[s] = ok ? pa[col] : mk(0.0, 0.0);
        fb0[s] = ok ? pb0[col] : mk(0.0, 0.0);
        fb1[s] = ok ? pb1[col] : mk(0.0, 0.0);
    }
    for (long long k0 = c0; k0 < c1; k0 += 4 * PF) {
#pragma unroll
        for (int s = 0; s < PF; s++) {
            const cplx wa = fa[s], wb0 = fb0[s], wb1 = fb1[s];
            const long long col = k0 + 4 * (PF + s) + t;          // same slot, PF steps ahead
            const bool ok = col < c1;
            fa[s] = ok ? pa[col] : mk(0.0, 0.0);
            fb0[s] = ok ? pb0[col] : mk(0.0, 0.0);
            fb1[s] = ok ? pb1[col] : mk(0.0, 0.0);
            // W_m conj(W_n): re = ar*br + ai*bi ; im = ai*br - ar*bi   (zero operands past c1 add nothing)
            dmma884(cr[0][0], cr[0][1], wa.x, wb0.x);
            dmma884(cr[0][0], cr[0][1], wa.y, wb0.y);
            dmma884(ci[0][0], ci[0][1], wa.y, wb0.x);
            dmma884(ci[0][0], ci[0][1], -wa.x, wb0.y);
            dmma884(cr[1][0], cr[1][1], wa.x, wb1.x);
            dmma884(cr[1][0], cr[1][1], wa.y, wb1.y);
            dmma884(ci[1][0], ci[1][1], wa.y, wb1.x);
            dmma884(ci[1][0], ci[1][1], -wa.x, wb1.y);
        }
    }
    double* Gp = G + ((long long)pair * gridDim.x + blockIdx.x) * PMAX * PMAX * 2;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int row = mt * 8 + g, col = (nt0 + j) * 8 + 2 * t;
        double* dst = Gp + ((long long)row * PMAX + col) * 2;
        *(double4*)dst = make_double4(cr[j][0], ci[j][0], cr[j][1], ci[j][1]);
    }
}

// Wext[rows] <- Q Wext[rows]: each warp owns 8-column strips, loads the 32x8 strip as B
// fragments, multiplies by Q (A fragments from padded shared planes) and stores in place.
__global__ void __launch_bounds__(NT, 2)
k_apply_mma(cplx* __restrict__ W, long long ldw, long long lenx, int chunk, PairSpec ps, int slot_base,
            const cplx* __restrict__ Q, const int* __restrict__ rotated, const int* __restrict__ done) {
    pdl_wait();
    pdl_trigger();
    if (done && *done) return;      // static (sync-free) mode: this SVD already converged
    const int pair = slot_base + blockIdx.y;
    if (!rotated[pair]) return;
    constexpr int QS = PMAX + 4;
    __shared__ double qr[PMAX * QS], qi[PMAX * QS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
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
            b[kt] = ok ? W[row * ldw + col] : mk(0.0, 0.0);
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
                bn[kt] = ok ? W[row * ldw + col] : mk(0.0, 0.0);
            }
        }
        double cr[4][2], ci[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) { cr[mt][0] = cr[mt][1] = ci[mt][0] = ci[mt][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < 8; kt++)
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                double ar = qr[(mt * 8 + g) * QS + kt * 4 + t];
                double ai = qi[(mt * 8 + g) * QS + kt * 4 + t];
                dmma884(cr[mt][0], cr[mt][1], ar, b[kt].x);
                dmma884(cr[mt][0], cr[mt][1], -ai, b[kt].y);
                dmma884(ci[mt][0], ci[mt][1], ar, b[kt].y);
                dmma884(ci[mt][0], ci[mt][1], ai, b[kt].x);
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
    cplx* W; double* G; cplx* Q; int* rotated; double* sig2; int* perm; int* notconv; cplx* T;
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
    static int eig_version = 3;        // QM_EIG=1 selects the first formulation for the 32-row pairs too (A/B runs)
    if (!eig_attr_set) {
        QM_CUDA(cudaFuncSetAttribute(k_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EIG_SMEM));
        QM_CUDA(cudaFuncSetAttribute(k_eig3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Eig3Smem)));
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
        tune_inner0 = e0 ? atoi(e0) : 2;
        tune_inner = e1 ? atoi(e1) : 1;
        tune_cross = e2 ? atoi(e2) : 1;
        tune_groups = e4 ? atoi(e4) : 1;
    }
    int sweeps = 0, converged = 0;
    const bool is_static = fixed_sweeps > 0;
    const int* donep = is_static ? w.notconv + 1 : nullptr;
    if (is_static) max_sweeps = fixed_sweeps;
    QM_CUDA(cudaMemsetAsync(w.notconv, 0, 4 * sizeof(int), st));
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
            if (eig_version == 3 && !g.single) {
                if (is_static) QM_LAUNCH(QM_CLS_SVD_EIG, s, k_eig3<<<np, NTE3, sizeof(Eig3Smem), s>>>(
                    w.G, eig_chunks, w.Q, tol2, max_inner, tune_ratio, cross_only, ps, slot, w.notconv, w.rotated,
                    w.sig2, donep));
                else QM_LAUNCH(QM_CLS_SVD_EIG, s, qm_launch_dep(k_eig3, dim3(np), dim3(NTE3), sizeof(Eig3Smem), s,
                    w.G, eig_chunks, w.Q, tol2, max_inner, tune_ratio, cross_only, ps, slot, w.notconv, w.rotated,
                    w.sig2, donep));
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
        if (grouped) {
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
        qm_prof_work(QM_CLS_SVD_GRAM, 8.0 * g.nrows * g.nrows * (double)g.len * g.npairs * g.rounds);
        qm_prof_work(QM_CLS_SVD_APPLY, 8.0 * g.nrows * g.nrows * (double)lenx * g.npairs * g.rounds);
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
