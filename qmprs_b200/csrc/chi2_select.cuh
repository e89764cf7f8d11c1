// chi=2 truncation bookkeeping for one bond, single thread (shared by mps_ops.cu and small_mps.cu).
//   S[4], Vh[4][4] (ld) : SVD of the padded (k x 4) matrix R_{i-1} T_i  (squared = 0), eigenvalues (squared = 1),
//   or Vh = the 4x4 Hermitian PSD matrix T^H L T itself (squared = 2; diagonalised here).
//   n = #{ S_j > cutoff*S_0 } capped at 2.  Canonical phase rule: each kept row of Vh is divided by the phase of its
//   first entry with |x|^2 >= (1-tie)*max|x|^2.  Csite[2][4] <- kept rows (zero padded); Vsel[4][2] <- their
//   conjugate transpose; bond[0] <- n; ambiguous[0] <- 1 when s_1 <= amb_rel * s_0 in a squared mode.
// Reference: compress(mode="right", max_bond=2) bookkeeping, qmprs/primitives/mps.py:881.
#pragma once
#include "common.cuh"
#include "small_linalg.cuh"

static __device__ void chi2_select_dev(const double* S, const cplx* Vh, long long ldvh, double cutoff, double tie,
                                       cplx* Csite, cplx* Vsel, int* bond, int squared, double amb_rel,
                                       int* ambiguous) {
    int n = 0;
    double sv[4];
    cplx vh_local[4][4];
    if (squared == 2) {
        // Vh points at the 4x4 Hermitian PSD matrix H = T^H L T itself: one-sided Jacobi on its columns
        // gives H V = U Sigma with Sigma = eigenvalues and V = eigenvectors; sort descending.
        cplx A[4][4], V[4][4];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) { A[i][j] = Vh[(long long)i * ldvh + j]; V[i][j] = mk(i == j ? 1.0 : 0.0, 0.0); }
        jacobi_cols(A, V, 4);
        double lam[4];
        int ord[4] = {0, 1, 2, 3};
        for (int j = 0; j < 4; j++) {
            double s2 = 0.0;
            for (int i = 0; i < 4; i++) s2 += cabs2(A[i][j]);
            lam[j] = sqrt(s2);
        }
        for (int a = 0; a < 3; a++)
            for (int b = a + 1; b < 4; b++)
                if (lam[ord[b]] > lam[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
        for (int j = 0; j < 4; j++) {
            sv[j] = sqrt(lam[ord[j]]);
            for (int c = 0; c < 4; c++) vh_local[j][c] = cconj(V[c][ord[j]]);
        }
    } else {
        for (int j = 0; j < 4; j++) {
            sv[j] = squared ? sqrt(S[j] > 0.0 ? S[j] : 0.0) : S[j];
            for (int c = 0; c < 4; c++) vh_local[j][c] = Vh[(long long)j * ldvh + c];
        }
    }
    if (squared && ambiguous && sv[1] <= amb_rel * sv[0]) ambiguous[0] = 1;
    double thr = cutoff * sv[0];
    for (int j = 0; j < 4; j++) n += (sv[j] > thr) ? 1 : 0;
    if (n < 1) n = 1;
    if (n > 2) n = 2;
    for (int j = 0; j < 2; j++) {
        cplx row[4];
        for (int c = 0; c < 4; c++) row[c] = (j < n) ? vh_local[j][c] : mk(0.0, 0.0);
        if (j < n) {
            double mx = 0.0;
            for (int c = 0; c < 4; c++) { double a = cabs2(row[c]); mx = a > mx ? a : mx; }
            int pick = 0;
            for (int c = 0; c < 4; c++) if (cabs2(row[c]) >= (1.0 - tie) * mx) { pick = c; break; }
            double a = sqrt(cabs2(row[pick]));
            cplx ph = (a > 0.0) ? mk(row[pick].x / a, row[pick].y / a) : mk(1.0, 0.0);
            for (int c = 0; c < 4; c++) row[c] = cmulc(row[c], ph);      // row / ph
        }
        for (int c = 0; c < 4; c++) {
            Csite[j * 4 + c] = row[c];
            Vsel[c * 2 + j] = cconj(row[c]);
        }
    }
    bond[0] = n;
}

