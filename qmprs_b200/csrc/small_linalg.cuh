// One-sided Jacobi on the columns of a d x d (d <= 4) complex matrix, single thread:
// A V = U Sigma on exit (columns of A orthogonal, V unitary).  Shared by the 4x4 polar factor of the
// sweeps (dense.cu) and the 4x4 Hermitian eigen-problem of the chi=2 truncation (mps_ops.cu).
#pragma once
#include "common.cuh"

static __device__ void jacobi_cols(cplx A[4][4], cplx V[4][4], int d) {
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
                // overflow-free form (see polar_conj_warp in dense.cu): no division by |g|
                double dd = 0.5 * (b - a);
                double hh = fma(dd, dd, mag2);
                double den = fabs(dd) + hh * rsqrt(hh);
                double R = rsqrt(den * den + mag2);
                double c = den * R, sR = copysign(R, dd);
                cplx se = mk(sR * g.x, sR * g.y), sec = cconj(se);               // s e^{+-i phi}
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
}

