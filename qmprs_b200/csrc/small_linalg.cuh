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
                double imag = rsqrt(mag2);
                double zeta = 0.5 * (b - a) * imag;
                double z1 = 1.0 + zeta * zeta;
                double t = copysign(1.0, zeta) / (fabs(zeta) + z1 * rsqrt(z1));
                double c = rsqrt(1.0 + t * t), s = c * t;
                cplx se = mk(s * g.x * imag, s * g.y * imag), sec = cconj(se);   // s e^{+-i phi}
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

