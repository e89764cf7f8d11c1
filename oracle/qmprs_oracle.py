"""CPU oracle for the qmprs MPS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``qmprs_b200`` / ``qmprs``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference (Qualition/qmprs) delegates every piece of arithmetic
to ``quimb==1.10.0`` (pyproject.toml:13), ``quick-core`` (pyproject.toml:12, unpinned
git dependency) and LAPACK through numpy/scipy.  None of quimb / quick / autoray /
qiskit is installed in this image and there is no network, so the reference cannot be
imported here and its tests hold no numeric golden vectors
(tests/synthesis/mps_encoding/test_sequential_encoding.py:67,89,116,121,150,155,181,207
are inequalities on unseeded inputs).  This file therefore RESTATES the published
algorithm of those dependencies (quimb 1.10.0: ``MatrixProductState.from_dense``,
``tensor_split`` / ``_trim_and_renorm_svd_result``, ``qr_stabilized``, ``left_canonize``,
``right_canonize``, ``right_compress`` / ``tensor_compress_bond``, ``gate_split``;
quick: ``_get_submps_indices``; scipy ``null_space``) and follows the reference's own
call sites line by line.  It is anchored on the reference's test inequalities and
on the outputs the real reference stack published: the README statistic (README.md:59-70),
the notebook's table of depths / U3 / CX counts for 2..12 qubits (30 integers, reproduced
exactly) and its fidelities for 7..12 qubits (each bracketed by two seeds of the same
distribution) -- tests/test_oracle.py, tests/test_transpile.py.  No fixture of the reference
pins a vector or a matrix, hence "unpinned" at that level.

Two gauge modes:

``verbatim``   numpy/scipy LAPACK calls exactly where the reference makes them
               (np.linalg.svd, np.linalg.qr, scipy.linalg.null_space).  Results carry
               LAPACK's arbitrary singular-vector phases, as the reference's do.
``canonical``  same algorithm plus a deterministic gauge: (i) after each chi=2
               truncation SVD every kept right-singular row is rotated so that its
               first entry of (near-)maximal magnitude is real positive, (ii) the
               isometry completion uses an explicit Householder-LQ with LAPACK
               zgelq2/zlarfg conventions (== scipy.linalg.null_space to 1e-15 on
               generic inputs) with a rounding-robust sign rule.  This mode is the
               parity target of the CUDA path (it is invariant to the phases an SVD
               implementation happens to return).

Array conventions: an MPS is a python list of N arrays ``A[i]`` of shape (l, 2, r);
site 0 is the most significant bit of the dense index (C-order reshapes), as in
quimb's ``to_dense`` / ``from_dense``.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg as sla

CUTOFF = 1e-10          # quimb default cutoff for tensor_split
TIE_REL = 1.0e-6        # canonical phase rule: "near-maximal" = |x|^2 >= (1-TIE_REL) max|x|^2
SIGN_TOL = 1.0e-12      # canonical Householder: Re(x0) >= -SIGN_TOL*||x|| counts as non-negative


# --------------------------------------------------------------------------------------
# quimb tensor_split semantics
# --------------------------------------------------------------------------------------
def trim(s, cutoff=CUTOFF, mode="rsum2", max_bond=None):
    """Number of singular values kept and renormalisation factor.

    Restates quimb 1.10.0 ``_compute_number_svals_to_keep`` /
    ``_trim_and_renorm_svd_result`` (decomp.py): mode ``rel`` keeps s_j > cutoff*s_0;
    mode ``rsum2`` discards the longest tail whose sum of squares is <= cutoff*sum(s^2)
    and (renorm default for the *sum* modes) rescales the kept values to preserve the
    Frobenius norm.  At least one value is kept.  Reached from mps.py:242 (rsum2),
    mps.py:451/453 (rel), mps.py:928-931/968-971 (rsum2).
    """
    s = np.asarray(s, dtype=np.float64)
    if mode == "rel":
        n = int(np.count_nonzero(s > cutoff * s[0]))
    elif mode == "rsum2":
        target = cutoff * float(np.sum(s * s))
        n = s.size
        ssum = 0.0
        for i in range(s.size - 1, -1, -1):
            ssum += float(s[i]) ** 2
            if ssum > target:
                break
            n -= 1
    else:
        raise ValueError(mode)
    n = max(n, 1)
    if max_bond:
        n = min(n, int(max_bond))
    f = 1.0
    if mode == "rsum2" and n < s.size:
        keep = float(np.sum(s[:n] ** 2))
        lose = float(np.sum(s[n:] ** 2))
        f = np.sqrt((keep + lose) / keep)
    return n, f


def qr_pos(x):
    """``qr_stabilized``: reduced QR with the diagonal of R made real non-negative."""
    q, r = np.linalg.qr(x)
    k = r.shape[0]
    d = np.diagonal(r)[:k].copy()
    ph = np.ones(k, dtype=np.complex128)
    nz = np.abs(d) > 0
    ph[nz] = d[nz] / np.abs(d[nz])
    q = q * ph[None, :]
    r = r * np.conj(ph)[:, None]
    return q, r


# --------------------------------------------------------------------------------------
# A1: statevector -> exact MPS          (mps.py:242 -> quimb MatrixProductState.from_dense)
# --------------------------------------------------------------------------------------
def from_dense(psi, n_sites, spectra=None):
    """Right-to-left TT-SVD, cutoff 1e-10 ``rsum2``, sqrt(s) absorbed on both sides.

    ``spectra`` (optional list) receives the full singular spectrum of every split,
    first entry = split at site N-1.
    """
    psi = np.asarray(psi, dtype=np.complex128).reshape(-1)
    N = int(n_sites)
    A = [None] * N
    T = psi.reshape(-1, 1)
    r = 1
    for i in range(N - 1, 0, -1):
        M = T.reshape(2 ** i, 2 * r)
        U, s, Vh = np.linalg.svd(M, full_matrices=False)
        if spectra is not None:
            spectra.append(s.copy())
        n, f = trim(s, CUTOFF, "rsum2")
        sq = np.sqrt(s[:n] * f)
        A[i] = (sq[:, None] * Vh[:n]).reshape(n, 2, r)
        T = U[:, :n] * sq[None, :]
        r = n
    A[0] = T.reshape(1, 2, r)
    return A


def to_dense(A):
    """Full contraction of an MPS to a 2^N vector (mps.py:270)."""
    x = A[0].reshape(-1, A[0].shape[2])
    for i in range(1, len(A)):
        l, _, r = A[i].shape
        x = (x @ A[i].reshape(l, 2 * r)).reshape(-1, r)
    return x.reshape(-1)


def mps_norm(A):
    """<psi|psi>^(1/2) by transfer matrices (mps.py:285)."""
    E = np.ones((1, 1), dtype=np.complex128)
    for a in A:
        l, _, r = a.shape
        t = (E @ a.reshape(l, 2 * r)).reshape(-1, r)            # (l'*2, r)
        E = np.conj(a.reshape(l * 2, r)).T @ t
    return float(np.sqrt(abs(E[0, 0])))


def bond_dims(A):
    return [a.shape[2] for a in A[:-1]]


# --------------------------------------------------------------------------------------
# canonical forms / compression       (mps.py:396-398, 451-453 -> quimb)
# --------------------------------------------------------------------------------------
def left_canon(A):
    """QR sweep left->right (quimb ``left_canonize``); returns a new list."""
    A = [a.copy() for a in A]
    for i in range(len(A) - 1):
        l, _, r = A[i].shape
        q, rr = qr_pos(A[i].reshape(l * 2, r))
        k = q.shape[1]
        A[i] = q.reshape(l, 2, k)
        r2 = A[i + 1].shape[2]
        A[i + 1] = (rr @ A[i + 1].reshape(r, 2 * r2)).reshape(k, 2, r2)
    return A


def right_canon(A, normalize=False):
    """LQ sweep right->left (quimb ``right_canonize``); norm ends on site 0."""
    A = [a.copy() for a in A]
    for i in range(len(A) - 1, 0, -1):
        l, _, r = A[i].shape
        q, rr = qr_pos(A[i].reshape(l, 2 * r).T)
        k = q.shape[1]
        A[i] = q.T.reshape(k, 2, r)
        l0 = A[i - 1].shape[0]
        A[i - 1] = (A[i - 1].reshape(l0 * 2, l) @ rr.T).reshape(l0, 2, k)
    if normalize:
        A[0] = A[0] / np.linalg.norm(A[0])
    return A


def canonical_row_phase(row):
    """Phase (unit complex) of the first entry of near-maximal magnitude of ``row``."""
    m2 = np.abs(row) ** 2
    j = int(np.argmax(m2 >= (1.0 - TIE_REL) * m2.max()))
    v = row[j]
    a = abs(v)
    return v / a if a > 0 else 1.0 + 0.0j


def right_compress(A, max_bond=None, gauge="verbatim", phase_fix=False, spectra=None):
    """Right->left truncation sweep on a LEFT-canonical MPS (quimb ``right_compress``
    -> ``tensor_compress_bond``: QR(T1), LQ(T2), SVD(R L) with cutoff 1e-10 ``rel``,
    ``absorb='left'``, no renormalisation).  ``phase_fix`` applies the canonical row
    phase rule (module docstring) after each SVD."""
    A = [a.copy() for a in A]
    for i in range(len(A) - 1, 0, -1):
        l0, _, b = A[i - 1].shape
        _, _, r = A[i].shape
        q1, r1 = qr_pos(A[i - 1].reshape(l0 * 2, b))
        q2t, l2t = qr_pos(A[i].reshape(b, 2 * r).T)
        q2, l2 = q2t.T, l2t.T
        U, s, Vh = np.linalg.svd(r1 @ l2, full_matrices=False)
        if spectra is not None:
            spectra.append(s.copy())
        n, _ = trim(s, CUTOFF, "rel", max_bond)
        left = q1 @ (U[:, :n] * s[None, :n])
        right = Vh[:n] @ q2
        if phase_fix:
            for j in range(n):
                ph = canonical_row_phase(right[j])
                right[j] = right[j] / ph
                left[:, j] = left[:, j] * ph
        A[i - 1] = left.reshape(l0, 2, n)
        A[i] = right.reshape(n, 2, r)
    return A


def compress_right(A, max_bond=None, gauge="verbatim", phase_fix=False, spectra=None):
    """``mps.compress(form='right'[, max_bond])`` (mps.py:451/453)."""
    return right_compress(left_canon(A), max_bond, gauge, phase_fix, spectra)


def build_mps(psi, n_sites, chi, spectra=None):
    """``MPS.from_statevector`` (mps.py:218-249): exact TT-SVD then truncation to chi.
    The reference's ``tensor_network_1d_compress`` ('dm' method) spans the same
    subspaces as a left-canonise + right->left truncated SVD sweep, and leaves the MPS
    right-canonical with the norm on site 0, not renormalised."""
    return compress_right(from_dense(psi, n_sites, spectra), max_bond=chi)


# --------------------------------------------------------------------------------------
# A5: isometry -> unitary completion   (mps.py:565-847)
# --------------------------------------------------------------------------------------
def null_space_householder(M):
    """Null-space basis of ``M`` (m x n, m < n, full row rank) by Householder LQ with
    LAPACK zgelq2/zlarfg conventions; columns of the result are orthonormal and
    satisfy M @ K = 0.  Sign rule made robust: beta = -||x|| when
    Re(x0) >= -SIGN_TOL*||x||, else +||x||."""
    M = np.array(M, dtype=np.complex128)
    m, n = M.shape
    Q = np.eye(n, dtype=np.complex128)          # accumulates H_0 H_1 ... (acting from the right)
    for i in range(m):
        x = np.conj(M[i, i:])                   # zlacgv
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0.0 and alpha.imag == 0.0:
            continue                            # tau = 0, H = I
        nrm = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm ** 2)
        beta = -nrm if alpha.real >= -SIGN_TOL * nrm else nrm
        tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
        v = np.empty(n - i, dtype=np.complex128)
        v[0] = 1.0
        v[1:] = x[1:] / (alpha - beta)
        # apply H = I - tau v v^H from the right to rows of M[:, i:] and to Q[:, i:]
        M[:, i:] = M[:, i:] - tau * np.outer(M[:, i:] @ v, np.conj(v))
        Q[:, i:] = Q[:, i:] - tau * np.outer(Q[:, i:] @ v, np.conj(v))
    # M_orig @ Q = [L 0]  ->  columns m.. of Q span the null space
    return Q[:, m:]


def null_space(M, gauge):
    if gauge == "verbatim":
        return sla.null_space(M)
    return null_space_householder(M)


def submps_indices(C):
    """quick ``_get_submps_indices`` (called at mps.py:802): maximal runs of sites
    joined by bonds of dimension >= 2."""
    N = len(C)
    out = []
    start = None
    for i in range(N):
        dl = C[i].shape[0]
        dr = C[i].shape[2]
        if dl < 2 and dr < 2:
            out.append((i, i))
        elif dl < 2 and dr >= 2:
            start = i
        elif dl >= 2 and dr < 2:
            out.append((start, i))
            start = None
    return out


def first_site_unitary(a, gauge):
    """mps.py:593-619.  ``a``: (1,2,2) tensor of the first site of a block."""
    K = null_space(np.conj(a.reshape(1, 4)), gauge)            # 4 x 3
    u = np.zeros((2, 2, 2, 2), dtype=np.complex128)
    u[0, 0] = a.reshape(2, 2)
    u[0, 1] = K[:, 0].reshape(2, 2)
    u[1, 0] = K[:, 1].reshape(2, 2)
    u[1, 1] = K[:, 2].reshape(2, 2)
    u = u.transpose(1, 0, 2, 3)
    return u.reshape(4, 4).T.copy()


def two_site_unitary(a, gauge):
    """mps.py:653-683.  ``a``: (2,2,2) tensor of an interior site of a block."""
    K = null_space(np.conj(a.reshape(2, 4)), gauge)            # 4 x 2
    if gauge == "verbatim":
        K = K / np.exp(1j * np.angle(K[0]))                    # mps.py:661-662
    else:
        # same rule, but a first entry that is zero up to rounding (structural zeros are
        # common after a disentangling layer) has no meaningful phase: leave that vector
        mag = np.abs(K[0])
        ph = np.where(mag > SIGN_TOL, K[0] / np.where(mag > 0, mag, 1.0), 1.0)
        K = K / ph
    u = np.zeros((2, 2, 2, 2), dtype=np.complex128)
    u[0] = a
    u[1] = K.reshape(2, 2, 2, 1).transpose(3, 2, 0, 1)
    u = u.transpose(1, 0, 2, 3)
    return u.reshape(4, 4).T.copy()


def last_site_unitary(a, single, gauge):
    """mps.py:723-740.  ``a``: (l,2,1) tensor; ``single`` when the block is one site."""
    if single:
        u = np.zeros((2, 2), dtype=np.complex128)
        u[0] = a.reshape(2)
        u[1] = null_space(np.conj(a.reshape(1, 2)), gauge).reshape(2)
    else:
        u = a.reshape(2, 2)
    return u.T.copy()


def is_unitary(G, atol=1e-8):
    return np.allclose(G @ np.conj(G).T, np.eye(G.shape[0]), atol=atol, rtol=1e-5)


def generate_unitary_layer(C, gauge):
    """mps.py:792-847.  Returns [(start, end, [G...])]."""
    layer = []
    for s, e in submps_indices(C):
        gates = []
        for i in range(s, e + 1):
            if i == e:
                gates.append(last_site_unitary(C[i], s == e, gauge))
            elif i == s:
                gates.append(first_site_unitary(C[i], gauge))
            else:
                gates.append(two_site_unitary(C[i], gauge))
        for g in gates:
            if not is_unitary(g):
                raise ValueError("All the generated unitaries must be unitary.")
        layer.append((s, e, gates))
    return layer


# --------------------------------------------------------------------------------------
# A4: chi=2 truncation                (mps.py:876-891)
# --------------------------------------------------------------------------------------
def chi2_truncate(B, gauge, spectra=None):
    """deepcopy + compress(mode='right', max_bond=2) + canonicalize('right', normalize)."""
    C = compress_right(B, max_bond=2, gauge=gauge, phase_fix=(gauge == "canonical"),
                       spectra=spectra)
    return right_canon(C, normalize=True)


# --------------------------------------------------------------------------------------
# A6: inverse layer application       (mps.py:944-971 -> quimb gate_ / gate_split_)
# --------------------------------------------------------------------------------------
def apply_inverse_layer(B, layer, spectra=None):
    """In place on the list ``B``.  Two-site gates: theta = G^H (A_i A_{i+1}), SVD with
    cutoff 1e-10 ``rsum2``, sqrt(s) to both sides, no max_bond."""
    for s, e, gates in layer:
        for i in range(e, s - 1, -1):
            g = np.conj(gates[i - s]).T
            if i == e:
                B[i] = np.einsum("op,lpr->lor", g, B[i])
            else:
                l, _, b = B[i].shape
                _, _, r = B[i + 1].shape
                x = (B[i].reshape(l * 2, b) @ B[i + 1].reshape(b, 2 * r)).reshape(l, 2, 2, r)
                th = np.einsum("abcd,lcdr->labr", g.reshape(2, 2, 2, 2), x).reshape(l * 2, 2 * r)
                U, sv, Vh = np.linalg.svd(th, full_matrices=False)
                if spectra is not None:
                    spectra.append(sv.copy())
                n, f = trim(sv, CUTOFF, "rsum2")
                sq = np.sqrt(sv[:n] * f)
                B[i] = (U[:, :n] * sq[None, :]).reshape(l, 2, n)
                B[i + 1] = (sq[:, None] * Vh[:n]).reshape(n, 2, r)
    return B


def zero_overlap(B):
    """``fidelity_with_zero_state`` (mps.py:1033-1039): conj(psi[0]); only element 0 of
    the dense vector is needed, so it is a product of the p=0 slices."""
    v = np.ones((1, 1), dtype=np.complex128)
    for a in B:
        v = v @ a[:, 0, :]
    return np.conj(v[0, 0])


# --------------------------------------------------------------------------------------
# dense statevector helpers (A8 / A9)
# --------------------------------------------------------------------------------------
def flatten_layers(layers):
    """Gates in application order: [(layer_idx, block_idx, tensor_idx, site, G)]."""
    out = []
    for li, layer in enumerate(layers):
        for bi, (s, e, gates) in enumerate(layer):
            for i in range(s, e + 1):
                out.append((li, bi, i - s, i, gates[i - s]))
    return out


def apply_gate_dense(c, n_sites, site, G):
    """out = G . in on axes (site[, site+1]) of the C-order reshape([2]*N)."""
    d = G.shape[0]
    k = 1 if d == 2 else 2
    L = 2 ** site
    R = 2 ** (n_sites - site - k)
    x = c.reshape(L, d, R)
    return np.einsum("ab,lbr->lar", G, x).reshape(-1)


def circuit_state(layers, n_sites):
    """A8: all gates applied in order to |0...0> (sequential.py:215-292, 443-447)."""
    c = np.zeros(2 ** n_sites, dtype=np.complex128)
    c[0] = 1.0
    for _, _, _, site, G in flatten_layers(layers):
        c = apply_gate_dense(c, n_sites, site, G)
    return c


NULL_REL = 1.0e-13      # canonical polar: singular values <= NULL_REL * s_max are null directions


def _polar_plain(E):
    u, _, vh = np.linalg.svd(E)
    return u @ vh


def polar_unitary(E, gauge="verbatim"):
    """u @ vh of the SVD of E (sequential.py:473-478).

    A rank-deficient E (a gate whose inputs do not span the full space: fresh |0> inputs,
    the left edge of a layer, block starts) leaves u @ vh undetermined on null(E); the
    reference gets whatever LAPACK returns there (decided by rounding noise), and that
    choice does feed back into later updates of the same sweep.  ``canonical`` fixes it
    independently of any basis: null(E) is mapped onto null(E^H) by the partial isometry
    closest to the identity, N_l polar(N_l^H N_r) N_r^H, the eps->0 limit of
    polar(E + eps*I)."""
    if gauge == "verbatim":
        return _polar_plain(E)
    u, s, vh = np.linalg.svd(E)
    smax = s[0] if s.size else 0.0
    keep = s > NULL_REL * smax if smax > 0 else np.zeros(s.shape, dtype=bool)
    r = int(np.count_nonzero(keep))
    d = E.shape[0]
    if r == d:
        return u @ vh
    P = u[:, :r] @ vh[:r]
    Nl = u[:, r:]                    # orthonormal basis of null(E^H)
    Nr = np.conj(vh[r:]).T           # orthonormal basis of null(E)
    X = _polar_plain(np.conj(Nl).T @ Nr)
    return P + Nl @ X @ np.conj(Nr).T


def sweep(target, layers, n_sites, gauge="verbatim"):
    """A9: one environment sweep (sequential.py:400-507).  ``target`` is the dense
    (un-normalised) chi-truncated state; ``layers`` is updated in place."""
    N = n_sites
    flat = flatten_layers(layers)
    c = circuit_state(layers, N)
    tbar = np.conj(target).copy()
    for li, bi, ti, site, G in reversed(flat):
        d = G.shape[0]
        k = 1 if d == 2 else 2
        L = 2 ** site
        R = 2 ** (N - site - k)
        c = apply_gate_dense(c, N, site, np.conj(G).T)                 # :460
        E = np.tensordot(tbar.reshape(L, d, R), c.reshape(L, d, R), axes=([0, 2], [0, 2]))  # :463
        Gn = np.conj(polar_unitary(E, gauge))                          # :473-491
        tbar = np.einsum("lor,ob->lbr", tbar.reshape(L, d, R), Gn).reshape(-1)   # :496
        layers[li][bi][2][ti] = Gn                                     # :501-505
    return layers


# --------------------------------------------------------------------------------------
# top level                           (base.py:96-104, sequential.py:330-398, 509-600)
# --------------------------------------------------------------------------------------
def prepare(psi, n_sites, chi, num_layers=1, num_sweeps=0, threshold=1 - 1e-6,
            gauge="canonical", record=None, schedule="DallOall"):
    """Restatement of ``Sequential.prepare_state``.  Returns a dict with

    ``layers``      list (application order) of [(start, end, [G...])],
    ``n_layers``    layers actually generated (early break, sequential.py:390),
    ``mps``         the chi-truncated MPS (right-canonical, norm on site 0),
    ``target``      its dense vector,
    ``overlaps``    conj(psi_k[0]) after each disentangling layer.
    ``record``      optional dict that receives per-stage spectra.
    """
    if not isinstance(num_layers, int) or num_layers < 1:
        raise ValueError("The number of layers must be a positive integer.")
    N = int(n_sites)
    psi = np.asarray(psi, dtype=np.complex128).reshape(-1)
    psi = psi / np.linalg.norm(psi)                                    # quick Ket
    rec = record if record is not None else {}
    rec.setdefault("tt_svd", [])
    rec.setdefault("truncate", [])
    rec.setdefault("gate_split", [])
    rec.setdefault("chi2", [])

    A = compress_right(from_dense(psi, N, rec["tt_svd"]), max_bond=chi, spectra=rec["truncate"])
    return prepare_mps(A, num_layers, num_sweeps, threshold, gauge, rec, schedule)


SCHEDULES = ("DallOall", "IterDiOall", "IterDiOi")


def prepare_mps(A, num_layers=1, num_sweeps=0, threshold=1 - 1e-6, gauge="canonical", record=None,
                schedule="DallOall"):
    """Restatement of ``Sequential.prepare_mps`` (sequential.py:588-600 -> :543-586) for an MPS given as
    site tensors (l, 2, r) in ANY gauge; same result dict as :func:`prepare`.

    ``schedule``: "DallOall" is what the reference runs (all layers by disentangling, then ``num_sweeps`` sweeps
    over all of them).  "IterDiOall" / "IterDiOi" are the two schedules the reference only names as future work
    (notebook cell at :459, docstring sequential.py:410, 428-432; Rudolph et al. 2022, the reference's [2]) --
    see :func:`prepare_mps_iterative`."""
    if not isinstance(num_layers, int) or num_layers < 1:
        raise ValueError("The number of layers must be a positive integer.")
    if schedule not in SCHEDULES:
        raise ValueError("`schedule` must be one of %s." % (SCHEDULES,))
    if schedule != "DallOall":
        return prepare_mps_iterative(A, num_layers, num_sweeps, threshold, gauge, record, schedule)
    rec = record if record is not None else {}
    rec.setdefault("gate_split", [])
    rec.setdefault("chi2", [])
    A = [np.asarray(a, dtype=np.complex128) for a in A]
    N = len(A)
    target = to_dense(A)                                               # sequential.py:440 (mps.mps, not normalised)

    # sequential.py:360-376
    B = [a.copy() for a in A]
    nrm = mps_norm(B)
    if not np.isclose(nrm, 1.0):
        B[-1] = B[-1] / nrm
    B = compress_right(B)
    B = right_canon(B, normalize=True)

    layers = []
    overlaps = []
    for _ in range(num_layers):
        sp6 = []
        sp4 = []
        C = chi2_truncate(B, gauge, sp4)                               # mps.py:878-887
        layer = generate_unitary_layer(C, gauge)                       # mps.py:889
        apply_inverse_layer(B, layer, sp6)                             # sequential.py:326
        rec["chi2"].append(sp4)
        rec["gate_split"].append(sp6)
        layers.append(layer)
        f = zero_overlap(B)
        overlaps.append(f)
        if np.isclose(f, 1 + 0j, atol=1 - threshold):                  # sequential.py:390
            break
    layers.reverse()                                                   # :396
    n_used = len(layers)

    for _ in range(num_sweeps):                                        # :532-539
        sweep(target, layers, N, gauge)

    return {"layers": layers, "n_layers": n_used, "mps": A, "target": target,
            "overlaps": overlaps, "n_sites": N}


def prepare_mps_iterative(A, num_layers, num_sweeps, threshold, gauge, record, schedule):
    """The two iterative schedules of Rudolph et al. 2022 that the reference lists as future work (notebook :459:
    "Iter DiOi, where we perform optimization on each layer as we generate them, or Iter DiOall, where we generate
    a layer and optimize that and its predecessors").  There is no reference code to follow; the building blocks
    are the reference's own (chi=2 truncation mps.py:849-891, inverse application :933-971, sweep
    sequential.py:400-507) composed as the paper describes:

    IterDiOi    layer k is generated from the residual |psi_k> (chi=2 truncation), then optimised ALONE for
                ``num_sweeps`` sweeps -- the rest of the circuit is fixed, so its environment is the one-layer
                circuit against |psi_k> -- and the optimised layer is taken out: |psi_{k+1}> = V_k^H |psi_k>.
    IterDiOall  layer k is generated from the residual, ALL layers generated so far are optimised for
                ``num_sweeps`` sweeps against the target, and the residual is rebuilt from the target with the
                optimised circuit: |psi_{k+1}> = V_k^H ... V_1^H |psi>.

    Same pre-conditioning, early break and result record as the default schedule; with ``num_sweeps = 0`` both
    reduce to it exactly."""
    rec = record if record is not None else {}
    rec.setdefault("gate_split", [])
    rec.setdefault("chi2", [])
    A = [np.asarray(a, dtype=np.complex128) for a in A]
    N = len(A)
    target = to_dense(A)
    B0 = [a.copy() for a in A]
    nrm = mps_norm(B0)
    if not np.isclose(nrm, 1.0):
        B0[-1] = B0[-1] / nrm
    B0 = right_canon(compress_right(B0), normalize=True)

    B = [a.copy() for a in B0]
    layers = []                                                        # application order: newest layer first
    overlaps = []
    for _ in range(num_layers):
        sp4, sp6 = [], []
        layer = generate_unitary_layer(chi2_truncate(B, gauge, sp4), gauge)
        rec["chi2"].append(sp4)
        if schedule == "IterDiOi":
            residual = to_dense(B)
            single = [layer]
            for _ in range(num_sweeps):
                sweep(residual, single, N, gauge)
            layers.insert(0, layer)
            apply_inverse_layer(B, layer, sp6)
        else:
            layers.insert(0, layer)
            for _ in range(num_sweeps):
                sweep(target, layers, N, gauge)
            if num_sweeps > 0:
                B = [a.copy() for a in B0]
                for lay in reversed(layers):                           # the layer applied last comes off first
                    sp6 = []
                    apply_inverse_layer(B, lay, sp6)
            else:
                apply_inverse_layer(B, layer, sp6)
        rec["gate_split"].append(sp6)
        f = zero_overlap(B)
        overlaps.append(f)
        if np.isclose(f, 1 + 0j, atol=1 - threshold):
            break
    return {"layers": layers, "n_layers": len(layers), "mps": A, "target": target,
            "overlaps": overlaps, "n_sites": N}


def emit_gates(layers, n_sites):
    """``_circuit_from_unitary_layers`` (sequential.py:155-213): (matrix, qubits) in
    emission order with the qubit reversal q = N-1-site."""
    N = n_sites
    out = []
    for layer in layers:
        for s, e, gates in layer:
            for i in range(s, e + 1):
                if i == e:
                    out.append((gates[i - s], abs(i - N + 1)))
                else:
                    out.append((gates[i - s], [abs(i - N + 2), abs(i - N + 1)]))
    return out


def simulate_emitted(gate_list, n_qubits):
    """Little-endian statevector simulation of ``emit_gates`` output (what quick's
    ``Circuit.get_statevector`` returns): qubit q is bit q of the index; for a two-qubit
    gate on [qa, qb] the matrix index is 2*bit(qb) + bit(qa)."""
    n = n_qubits
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1.0
    for G, q in gate_list:
        t = psi.reshape([2] * n)                    # axis a <-> qubit n-1-a
        if isinstance(q, (list, tuple)):
            qa, qb = q
            axb, axa = n - 1 - qb, n - 1 - qa
            g = np.asarray(G).reshape(2, 2, 2, 2)   # [ob, oa, ib, ia]
            t = np.tensordot(g, t, axes=([2, 3], [axb, axa]))
            t = np.moveaxis(t, [0, 1], [axb, axa])
        else:
            ax = n - 1 - q
            t = np.tensordot(np.asarray(G), t, axes=([1], [ax]))
            t = np.moveaxis(t, 0, ax)
        psi = np.ascontiguousarray(t).reshape(-1)
    return psi


def circuit_fidelity(psi, layers, n_sites):
    """|<psi|circuit|0..0>| with psi normalised (README.md:66)."""
    psi = np.asarray(psi, dtype=np.complex128).reshape(-1)
    psi = psi / np.linalg.norm(psi)
    return float(abs(np.vdot(psi, circuit_state(layers, n_sites))))


def count_gates(layers):
    n2 = sum(1 for *_, g in flatten_layers(layers) if g.shape[0] == 4)
    n1 = sum(1 for *_, g in flatten_layers(layers) if g.shape[0] == 2)
    return n2, n1


def random_state(n_qubits, seed):
    """Reference input distribution (README.md:52-53; test_sequential_encoding.py:42-43)."""
    rng = np.random.default_rng(seed)
    v = rng.random(2 ** n_qubits) + 1j * rng.random(2 ** n_qubits)
    return v / np.linalg.norm(v)
