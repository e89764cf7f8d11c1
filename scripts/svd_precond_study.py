"""Numpy study (CPU, no GPU code): mixed-precision pre-conditioning of the one-sided Jacobi SVD.

Question for the next round: the 1024-row SVDs of the inverse-layer step need 12-14 FP64 sweeps, ~9 of them in the
slow (linear) phase.  If that phase runs in complex64 (TF32 tensor cores, half the bytes) and only hands its
accumulated rotation V to the FP64 solver, how many FP64 sweeps are left and is the spectrum still exact?

  phase 1  one-sided Jacobi on W.astype(complex64), rotations accumulated in complex64, stop at |cos| < stop32
  glue     V <- V (3 I - V^H V) / 2 in FP64 (Newton-Schulz; V32 is unitary only to 1e-6), W1 = V W in FP64
  phase 2  FP64 one-sided Jacobi on W1 to 1e-14

Result (this container, uniform random matrices, 64/128/256 rows): the complex64 phase needs as many sweeps as the
FP64 slow phase (9-12, down to |cos| ~ 1e-3..1e-4), after which 3-4 FP64 sweeps remain (plus the verification sweep
the device code skips); two Newton-Schulz steps restore unitarity to 7e-16 and the spectrum stays exact (1e-15 of
s_max, 3e-14 relative over the kept range).  So the gain is bounded by the cost ratio of a complex64 sweep to an FP64
sweep: ~10 x r + 3.5 sweeps against 12-13, i.e. 1.4x for r = 0.5 -- not the 3-4x a direct (non-Jacobi) complex64
eigen-solver of the Gram matrix would give (np.linalg.eigh in complex64 + 3 FP64 sweeps in the same experiment).

usage: python scripts/svd_precond_study.py [n ...]
"""
import sys
import time

import numpy as np


def rr_rounds(n):
    idx = list(range(n))
    out = []
    for _ in range(n - 1):
        out.append((np.array([min(idx[i], idx[n - 1 - i]) for i in range(n // 2)]),
                    np.array([max(idx[i], idx[n - 1 - i]) for i in range(n // 2)])))
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return out


def jacobi(W, tol, acc=None, max_sweeps=40):
    """Row-orthogonalising cyclic Jacobi; returns (W, acc, sweeps, history of the largest |cos| per sweep)."""
    W = W.copy()
    n = W.shape[0]
    rdt = W.real.dtype
    rounds = rr_rounds(n)
    hist = []
    for s in range(max_sweeps):
        mx = 0.0
        for p, q in rounds:
            Wp, Wq = W[p], W[q]
            a = np.einsum("ij,ij->i", Wp.conj(), Wp).real
            b = np.einsum("ij,ij->i", Wq.conj(), Wq).real
            g = np.einsum("ij,ij->i", Wp, Wq.conj())
            m2 = (g.real ** 2 + g.imag ** 2).astype(rdt)
            rel = np.sqrt(m2 / np.maximum(a * b, np.finfo(rdt).tiny))
            mx = max(mx, float(rel.max()))
            act = rel > tol
            d = 0.5 * (b - a)
            den = np.abs(d) + np.sqrt(d * d + m2)
            R = 1.0 / np.sqrt(den * den + m2 + np.finfo(rdt).tiny)
            c = np.where(act, den * R, 1.0).astype(rdt)
            off = (np.where(act, -np.copysign(R, d), 0.0).astype(rdt) * g).astype(W.dtype)
            W[p] = c[:, None] * Wp + off[:, None] * Wq
            W[q] = -off.conj()[:, None] * Wp + c[:, None] * Wq
            if acc is not None:
                Ap, Aq = acc[p], acc[q]
                acc[p] = c[:, None] * Ap + off[:, None] * Aq
                acc[q] = -off.conj()[:, None] * Ap + c[:, None] * Aq
        hist.append(mx)
        if mx <= tol:
            break
    return W, acc, s + 1, hist


def study(n, stop32=1e-3, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.random((n, n)) + 1j * rng.random((n, n))          # reference input distribution (one dominant value)
    sref = np.linalg.svd(A, compute_uv=False)
    t0 = time.time()
    _, _, s64, h64 = jacobi(A, 1e-14)
    W32, V32, s32, h32 = jacobi(A.astype(np.complex64), stop32, acc=np.eye(n, dtype=np.complex64))
    V = V32.astype(np.complex128)
    for _ in range(2):
        V = 0.5 * (3.0 * V - (V @ V.conj().T) @ V)             # rows stay an orthonormal frame: V V^H -> I
    unit = np.abs(V @ V.conj().T - np.eye(n)).max()
    W1 = V @ A
    W2, _, s2, h2 = jacobi(W1, 1e-14)
    s = np.sort(np.linalg.norm(W2, axis=1))[::-1]
    err = np.abs(s - sref).max() / sref[0]
    big = sref > 1e-5 * sref[0]
    rel = (np.abs(s - sref)[big] / sref[big]).max()
    print(f"n={n:5d}  plain fp64: {s64} sweeps | fp32 phase: {s32} sweeps to {stop32:g} (last {h32[-1]:.1e}) | "
          f"|VV^H-I| after 2 Newton-Schulz {unit:.1e} | fp64 phase: {s2} sweeps {['%.0e' % x for x in h2]} | "
          f"spectrum max err/s0 {err:.1e}, max rel (kept range) {rel:.1e} | {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    for n in [int(x) for x in sys.argv[1:]] or [64, 128, 256]:
        study(n)
        study(n, stop32=1e-4)
