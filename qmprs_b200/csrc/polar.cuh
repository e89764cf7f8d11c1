// 4x4 / 2x2 polar factor of the environment tensors (sequential.py:473-478) and the warp transpose-reduce
// shared by the streaming environment kernels (dense.cu) and the shared-memory small-state kernel
// (dense_small.cu).  Header-only so that each translation unit gets its own copies: ptxas derives the register
// budget of the out-of-line routines from the whole call graph of a unit, and mixing the 254-register streaming
// kernels with the small-state kernel in one unit capped the former at 128 registers (1.1 KB of spills,
// 30 -> 45 us per gate-step at 20 qubits).
#pragma once
#include "common.cuh"
#include "small_linalg.cuh"

namespace {

// ---------------------------------------------------------------------------------
// Polar factor of a d x d matrix (d <= 4) by one-sided Jacobi:  E V = U Sigma, P = U V^H.
// Rank-deficient E (gates whose inputs do not span the full space, e.g. a fresh |0> input
// or the left edge of a layer) leaves P undetermined on null(E); the reference inherits
// whatever LAPACK returns there.  Canonical rule (oracle `canonical` mode): directions with
// sigma <= 1e-13 sigma_max are null, and null(E) is mapped onto null(E^H) by the partial
// isometry closest to the identity, N_l polar(N_l^H N_r) N_r^H (the eps->0 limit of
// polar(E + eps I)); it depends on E only, not on any basis choice.
// Single thread.  polar_conj writes conj(P) (sequential.py:478-491).
// ---------------------------------------------------------------------------------
// Gram-Schmidt completion: for every column j with isnull[j], pick the standard basis vector
// with the largest residual against all fixed columns, orthogonalise twice, normalise.
__device__ void complete_columns(cplx U[4][4], const bool* isnull, int d) {
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
                    for (int i = 0; i < d; i++) ccfma(dot, U[i][c], v[i]);
                    for (int i = 0; i < d; i++) v[i] = csub(v[i], cmul(U[i][c], dot));
                }
            double nr = 0.0;
            for (int i = 0; i < d; i++) nr += cabs2(v[i]);
            if (nr > best * (1.0 + 1e-9)) {
                best = nr;
                for (int i = 0; i < d; i++) bestv[i] = v[i];
            }
        }
        double inv = rsqrt(best);
        for (int i = 0; i < d; i++) U[i][j] = cscale(bestv[i], inv);
        fixed[j] = true;
    }
}

// M (d x d) <- its polar factor.  NESTED: canonical completion of the null directions.
template <bool NESTED>
__device__ void polar_factor(cplx M[4][4], int d) {
    cplx A[4][4], V[4][4];
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) { A[i][j] = M[i][j]; V[i][j] = mk(i == j ? 1.0 : 0.0, 0.0); }
    jacobi_cols(A, V, d);
    double sig[4], smax = 0.0;
    for (int j = 0; j < d; j++) {
        double s = 0.0;
        for (int i = 0; i < d; i++) s += cabs2(A[i][j]);
        sig[j] = sqrt(s);
        smax = sig[j] > smax ? sig[j] : smax;
    }
    bool isnull[4];
    int m = 0;
    for (int j = 0; j < d; j++) {
        isnull[j] = !(sig[j] > 1e-13 * smax) || smax == 0.0;
        if (isnull[j]) m++;
        else {
            double inv = 1.0 / sig[j];
            for (int i = 0; i < d; i++) A[i][j] = cscale(A[i][j], inv);
        }
    }
    if (m > 0) {
        complete_columns(A, isnull, d);                  // columns isnull[] of A: some basis N_l of null(M^H)
        if (NESTED) {
            int nidx[4];
            int c = 0;
            for (int j = 0; j < d; j++) if (isnull[j]) nidx[c++] = j;
            cplx X[4][4];                                // X = N_l^H N_r  (m x m), N_r = V[:, null]
            for (int a = 0; a < m; a++)
                for (int b = 0; b < m; b++) {
                    cplx sacc = mk(0.0, 0.0);
                    for (int i = 0; i < d; i++) ccfma(sacc, A[i][nidx[a]], V[i][nidx[b]]);
                    X[a][b] = sacc;
                }
            polar_factor<false>(X, m);
            cplx Nl[4][4];
            for (int i = 0; i < d; i++)
                for (int a = 0; a < m; a++) Nl[i][a] = A[i][nidx[a]];
            for (int i = 0; i < d; i++)
                for (int b = 0; b < m; b++) {
                    cplx sacc = mk(0.0, 0.0);
                    for (int a = 0; a < m; a++) cfma(sacc, Nl[i][a], X[a][b]);
                    A[i][nidx[b]] = sacc;
                }
        }
    }
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            cplx sacc = mk(0.0, 0.0);
            for (int k = 0; k < d; k++) cfmac(sacc, A[i][k], V[j][k]);      // U V^H
            M[i][j] = sacc;
        }
}

__device__ __noinline__ void polar_conj(const cplx* E, int d, cplx* out) {
    cplx M[4][4];
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) M[i][j] = E[i * d + j];
    polar_factor<true>(M, d);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) out[i * d + j] = cconj(M[i][j]);
}

// ---------------------------------------------------------------------------------
// Warp-cooperative version of polar_conj: lanes 0-15 hold A[i][j] (i = lane/4 % 4, j = lane%4),
// lanes 16-31 hold V[i][j].  The three perfect matchings of the four columns are the xor
// patterns 1,2,3, so the partner column of a lane is lane^m; Gram sums over the rows are
// xor-4 / xor-8 shuffles; the two disjoint rotations of a round run concurrently.  A 2x2
// input is embedded as diag(E, I).  Rank-deficient inputs (rare: 1 + L of the N*L gates)
// fall back to the single-thread routine so that the canonical completion stays in one place.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ cplx shfl_xor_c(cplx v, int m) {
    return mk(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// The Jacobi rounds on the 32 lanes' elements (x: A[i][j] in lanes 0-15, V[i][j] in lanes 16-31).  A sweep is the
// three perfect matchings.  The loop ends after a sweep in which every rotation was tiny (|cos|^2 <= 1e-16 between
// the two columns): one-sided Jacobi converges quadratically, so such a sweep leaves cosines of ~1e-16 -- a
// verification sweep (three more rounds of the serial chain, ~1.2 us) would find nothing.  Returns the rounds done.
__device__ __forceinline__ int polar_jacobi_rounds(cplx& x, int half, int j) {
    const double tol2 = 4e-30;
    unsigned big = 0;                                // a rotation above that level happened in the current sweep
    int rounds = 0;
    for (int it = 0; it < 90; it++) {
        rounds++;
        const int m = (it % 3) + 1;
        cplx y = shfl_xor_c(x, m);                   // partner column, same row, same matrix
        const bool isp = j < (j ^ m);
        double na = cabs2(x), nb = cabs2(y);
        cplx gg = isp ? ccmul(x, y) : ccmul(y, x);   // conj(a_p) a_q contribution of this row
        na += __shfl_xor_sync(0xffffffffu, na, 4); na += __shfl_xor_sync(0xffffffffu, na, 8);
        nb += __shfl_xor_sync(0xffffffffu, nb, 4); nb += __shfl_xor_sync(0xffffffffu, nb, 8);
        gg = cadd(gg, shfl_xor_c(gg, 4)); gg = cadd(gg, shfl_xor_c(gg, 8));
        // the V half takes the sums of the A half
        double na2 = __shfl_xor_sync(0xffffffffu, na, 16), nb2 = __shfl_xor_sync(0xffffffffu, nb, 16);
        cplx gg2 = shfl_xor_c(gg, 16);
        if (half) { na = na2; nb = nb2; gg = gg2; }
        const double a = isp ? na : nb, b = isp ? nb : na;      // |a_p|^2, |a_q|^2
        const double mag2 = cabs2(gg);
        const bool rot = (a > 0.0 && b > 0.0 && mag2 > tol2 * a * b);
        if (rot) {
            // overflow-free form (no 1/|g|: a column that is numerically null shrinks geometrically under the
            // rotations, |g|^2 reaches the denormals and zeta^2 = ((b-a)/2|g|)^2 overflowed to inf -> NaN):
            // dd = (b-a)/2, h = sqrt(dd^2+|g|^2), den = |dd|+h, R = 1/sqrt(den^2+|g|^2): c = den R, s e^{i phi} = sign(dd) R g
            const double dd = 0.5 * (b - a);
            const double hh = fma(dd, dd, mag2);                 // > 0 since mag2 > 0
            const double den = fabs(dd) + hh * rsqrt(hh);       // (sqrt() would pull in a slow-path call and cap registers)
            const double R = rsqrt(den * den + mag2);
            const double c = den * R, sR = copysign(R, dd);
            const cplx se = mk(sR * gg.x, sR * gg.y);           // s e^{i phi}
            // x_p' = c x_p - s e^{-i phi} x_q ;  x_q' = s e^{i phi} x_p + c x_q
            if (isp) x = csub(cscale(x, c), cmul(cconj(se), y));
            else x = cadd(cmul(se, y), cscale(x, c));
        }
        big |= __ballot_sync(0xffffffffu, rot && mag2 > 1e-16 * a * b);
        if (m == 3) {                                // end of a sweep
            if (!big) break;
            big = 0;
        }
    }
    return rounds;
}

// vwarm (optional, global, 16 cplx): right singular vectors found for this gate in the previous sweep.
// The environments change little from sweep to sweep, so E V_prev already has nearly orthogonal
// columns and the Jacobi iteration starts in its quadratic regime (2 sweeps instead of 5-6).
// vwarm_out: where the vectors found now are stored (the same array in the per-gate kernels; a second buffer in the
// persistent kernel, where every CTA reads the warm start while CTA 0 writes the new one; NULL = do not store).
//
// Rank-deficient E (every gate of the first layer, whose inputs include a fresh |0>: 1 gate-step in 10 of a
// 12-qubit / 10-layer circuit) stays in the warp: with U_r, V_r the singular vectors of the non-null part,
// P_l = I - U_r U_r^H and N_r the null columns of V, the canonical completion  N_l polar(N_l^H N_r) N_r^H  is the
// partial isometry of P_l N_r N_r^H, whose row and column spaces are orthogonal to those of E.  So the null columns of
// A = E V are REPLACED by sigma_max P_l V_j and the same Jacobi iteration continues: it orthogonalises those columns
// among themselves (the others are already orthogonal to them), and U V^H is the canonical polar factor.  Only when
// the replacement is itself rank-deficient (N_l^H N_r singular; E = 0) the single-thread routine takes over.
// Round 1 sent every rank-deficient gate there: ~120 k cycles each, half of the polar time of a small state.
// Returns the number of Jacobi rounds done, negative when the single-thread routine ran (instrumentation only).
__device__ int polar_conj_warp(const cplx* Es, int d, cplx* gate_out, cplx* scratch /* smem, 32 cplx */,
                               const cplx* vwarm, cplx* vwarm_out) {
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4, i = (lane >> 2) & 3, j = lane & 3;
    cplx x;
    cplx vin = mk(i == j ? 1.0 : 0.0, 0.0);
    if (vwarm) {
        cplx v = vwarm[i * 4 + j];
        if (__ballot_sync(0xffffffffu, cabs2(v) > 0.0)) vin = v;     // all-zero = no warm start yet
    }
    scratch[lane] = (half == 0) ? ((i < d && j < d) ? Es[i * d + j] : mk(i == j ? 1.0 : 0.0, 0.0)) : vin;
    __syncwarp();
    if (half == 0) {                                  // E V_prev, the four products summed as a tree
        const cplx p0 = cmul(scratch[i * 4 + 0], scratch[16 + 0 * 4 + j]), p1 = cmul(scratch[i * 4 + 1], scratch[16 + 1 * 4 + j]);
        const cplx p2 = cmul(scratch[i * 4 + 2], scratch[16 + 2 * 4 + j]), p3 = cmul(scratch[i * 4 + 3], scratch[16 + 3 * 4 + j]);
        x = cadd(cadd(p0, p1), cadd(p2, p3));
    } else {
        x = vin;
    }
    __syncwarp();
    int rounds = polar_jacobi_rounds(x, half, j);
    // column norms of A, null detection (sigma_j <= 1e-13 sigma_max)
    double n2 = cabs2(x);
    n2 += __shfl_xor_sync(0xffffffffu, n2, 4); n2 += __shfl_xor_sync(0xffffffffu, n2, 8);
    double nmax = n2;
    nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, 1));
    nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, 2));
    bool isnull = (half == 0) && (!(n2 > 1e-26 * nmax) || nmax == 0.0);
    unsigned nullcols = __ballot_sync(0xffffffffu, isnull) & 0xFu;          // bit j: column j of A is null
    if (nullcols != 0u && nullcols != 0xFu) {
        const double amax = __shfl_sync(0xffffffffu, nmax, 0);               // sigma_max^2 (of the A half)
        const bool nj = (nullcols >> j) & 1u;
        scratch[lane] = (half == 0) ? (nj ? mk(0.0, 0.0) : cscale(x, rsqrt(n2))) : x;   // [0..15] = U_r (0 in null columns), [16..31] = V
        __syncwarp();
        // C[k][j] = sum_r conj(U[r][k]) V[r][j], held by lane (k = i, j) of both halves
        cplx cc = mk(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < 4; r++) ccfma(cc, scratch[r * 4 + i], scratch[16 + r * 4 + j]);
        // W[i][j] = V[i][j] - sum_k U[i][k] C[k][j]  =  (P_l V)[i][j]
        cplx w = scratch[16 + i * 4 + j];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const cplx ckj = mk(__shfl_sync(0xffffffffu, cc.x, k * 4 + j), __shfl_sync(0xffffffffu, cc.y, k * 4 + j));
            w = csub(w, cmul(scratch[i * 4 + k], ckj));
        }
        if (half == 0 && nj) x = cscale(w, amax * rsqrt(amax));
        __syncwarp();
        rounds += polar_jacobi_rounds(x, half, j);
        n2 = cabs2(x);
        n2 += __shfl_xor_sync(0xffffffffu, n2, 4); n2 += __shfl_xor_sync(0xffffffffu, n2, 8);
        nmax = n2;
        nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, 1));
        nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, 2));
        isnull = (half == 0) && (!(n2 > 1e-26 * nmax) || nmax == 0.0);
        nullcols = __ballot_sync(0xffffffffu, isnull) & 0xFu;
    }
    if (nullcols) {
        if (lane == 0) {
            cplx E[16], P[16];
            for (int k = 0; k < d * d; k++) E[k] = Es[k];
            polar_conj(E, d, P);
            for (int k = 0; k < d * d; k++) gate_out[k] = P[k];
        }
        if (vwarm_out && half) vwarm_out[i * 4 + j] = mk(0.0, 0.0);
        return -rounds;
    }
    if (vwarm_out && half) vwarm_out[i * 4 + j] = x;
    if (half == 0) x = cscale(x, rsqrt(n2));
    scratch[lane] = x;                                // [0..15] = U, [16..31] = V
    __syncwarp();
    if (lane < 16 && i < d && j < d) {                // U V^H, the four products summed as a tree
        const cplx p0 = cmulc(scratch[i * 4 + 0], scratch[16 + j * 4 + 0]), p1 = cmulc(scratch[i * 4 + 1], scratch[16 + j * 4 + 1]);
        const cplx p2 = cmulc(scratch[i * 4 + 2], scratch[16 + j * 4 + 2]), p3 = cmulc(scratch[i * 4 + 3], scratch[16 + j * 4 + 3]);
        gate_out[i * d + j] = cconj(cadd(cadd(p0, p1), cadd(p2, p3)));
    }
    return rounds;
}


// warp transpose-reduce of 32 doubles: lane L ends with the warp sum of v[L] in v[0]
__device__ __forceinline__ double warp_reduce32(double* v, int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            double send = upper ? v[i] : v[i + n / 2];
            double keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

}  // namespace
