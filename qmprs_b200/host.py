"""Host orchestration of the MPS hot path on top of the C-ABI kernels.

Every function mirrors one row of SURVEY.md section 8(a) and cites the reference lines
it replaces (paths relative to the reference repo).  An MPS is a python list of N
device tensors ``A[i]`` of shape (l, 2, r) (quimb's 'lpr' layout, edge bonds of size 1);
site 0 is the most significant bit of the dense index.

``K`` is the kernel handle (:class:`qmprs_b200.kernels.CudaKernels`).  All arithmetic is
done by its kernels; this module only sequences launches, reshapes views and reads the
few scalars that decide shapes (ranks, block structure, the early-break overlap).
"""
from __future__ import annotations

import numpy as np

from .kernels import CUTOFF, MODE_REL, MODE_RSUM2


# --------------------------------------------------------------------------------------
# A1  statevector -> exact MPS        qmprs/primitives/mps.py:242 (quimb from_dense)
# --------------------------------------------------------------------------------------
def from_dense(K, psi, n_sites, spectra=None):
    """Right-to-left TT-SVD, cutoff 1e-10 'rsum2', sqrt(s) absorbed on both sides."""
    N = int(n_sites)
    A = [None] * N
    T = psi.reshape(-1, 1)
    r = 1
    for i in range(N - 1, 0, -1):
        M = T.reshape(2 ** i, 2 * r)
        U, S, Vh = K.svd(M, backmult=True)
        k = S.shape[0]
        if spectra is not None:
            spectra.append(S)
        rank, f = K.trim(S, k, CUTOFF, MODE_RSUM2)
        n = K.read_int(rank, expect=k)
        A[i] = K.scale_copy(Vh[:n], S, f, mode=1, half_power=True).reshape(n, 2, r)
        T = K.scale_copy(U[:, :n], S, f, mode=2, half_power=True)
        r = n
    A[0] = T.reshape(1, 2, r)
    return A


def to_dense(K, A):
    """Full contraction to a 2^N vector (mps.py:270)."""
    x = A[0].reshape(-1, A[0].shape[2])
    for i in range(1, len(A)):
        l, _, r = A[i].shape
        x = K.gemm(x, A[i].reshape(l, 2 * r)).reshape(-1, r)
    return x.reshape(-1)


def bond_dims(A):
    return [int(a.shape[2]) for a in A[:-1]]


def _fused_small(K, A):
    """Static (CUDA-graph) mode on a register whose bonds all fit the fused small-register kernels
    (csrc/small_mps.cu): a batch of such states is bound by the number of kernel nodes, not by arithmetic."""
    return (getattr(K, "static", False) and getattr(K, "fused_small", True) and hasattr(K, "split_absorb")
            and max(int(a.shape[0]) for a in A) <= K.FUSED_MAX_BOND and max(int(a.shape[2]) for a in A) <= K.FUSED_MAX_BOND)


def copy_mps(K, A):
    return [K.scale_copy(a.reshape(a.shape[0] * 2, a.shape[2])).reshape(a.shape) for a in A]


# --------------------------------------------------------------------------------------
# A2/A3  canonical forms and truncation   mps.py:247, :396-398, :451-453
# --------------------------------------------------------------------------------------
def left_canon(K, A):
    """QR sweep left->right with non-negative diag(R) (quimb left_canonize)."""
    A = list(A)
    for i in range(len(A) - 1):
        l, _, r = A[i].shape
        Q, R = K.qr(A[i].reshape(l * 2, r))
        k = Q.shape[1]
        A[i] = Q.reshape(l, 2, k)
        r2 = A[i + 1].shape[2]
        A[i + 1] = K.gemm(R, A[i + 1].reshape(r, 2 * r2)).reshape(k, 2, r2)
    return A


def right_compress(K, A, max_bond=None, spectra=None, cutoff=CUTOFF):
    """Right->left truncation sweep on a LEFT-canonical MPS: SVD of each site matrix,
    cutoff 1e-10 'rel' (+ max_bond), singular values absorbed to the left, no renorm
    (quimb right_compress / tensor_compress_bond; the QR/LQ reduction of the reference
    is an identity on a left-canonical input)."""
    A = list(A)
    for i in range(len(A) - 1, 0, -1):
        b, _, r = A[i].shape
        l0 = A[i - 1].shape[0]
        U, S, Vh = K.svd(A[i].reshape(b, 2 * r), backmult=True)
        k = S.shape[0]
        if spectra is not None:
            spectra.append(S)
        rank, _ = K.trim(S, k, cutoff, MODE_REL, max_bond or 0)
        n = K.read_int(rank, expect=min(k, max_bond) if max_bond else k)
        US = K.scale_copy(U[:, :n], S, None, mode=2, half_power=False)
        A[i] = K.scale_copy(Vh[:n]).reshape(n, 2, r)
        A[i - 1] = K.gemm(A[i - 1].reshape(l0 * 2, b), US).reshape(l0, 2, n)
    return A


def canonicalize_truncate(K, A, max_bond=None, spectra=None, cutoff=CUTOFF):
    """``tensor_network_1d_compress(max_bond)`` (mps.py:247) and ``compress(form='right')``
    (mps.py:451/453): right-canonical MPS with the norm on site 0."""
    return right_compress(K, left_canon(K, A), max_bond, spectra, cutoff)


def mirror(K, A):
    """Site-reversed MPS: tensors (l,2,r) -> (r,2,l) in reverse order (used to run the
    left-handed variants of canonicalize/compress through the right-handed kernels)."""
    return [K.reverse3(a) for a in reversed(A)]


def normalize_site0(K, A):
    """``right_canonize(normalize=True)`` on an already right-canonical MPS (mps.py:398)."""
    a0 = A[0]
    K.div_sqrt(a0, K.vdot(a0, a0))
    return A


# --------------------------------------------------------------------------------------
# A4 + A5  chi=2 truncation and unitary completion   mps.py:849-891, :565-847
# --------------------------------------------------------------------------------------
def chi2_layer(K, B, debug=None, accurate=False):
    """Returns (gates [N,16] device, kinds list[int] host).  ``B`` is not modified
    (the reference works on a deepcopy, mps.py:878).

    The reference left-canonises the copy and then truncates bond by bond from the right
    (QR, LQ, SVD of R.L keeping <= 2 values).  Every truncation step only needs the two
    leading right singular vectors of  R_{i-1} T_i  (k x 4), T_i being site i contracted
    with the already truncated right part and R_{i-1} any square root of the left
    environment L_{i-1} = sum_p B^H L B.  Bonds are zero-padded to 2 so that all shapes are
    static and the bond dimensions stay on the device until the block structure is read.

    Fast path: L_i by two DMMA GEMMs per site and the 4x4 Hermitian problem T^H L T.  It
    squares the singular values, so when s_1 <= 1e-3 s_0 anywhere (rank decision at the
    1e-10 cutoff or second vector not resolvable: product-like / structured states) the
    layer is redone with the ``accurate`` path: R factors from Householder QR and the
    SVD of R_{i-1} T_i itself.
    """
    N = len(B)
    fused = not accurate and _fused_small(K, B)
    facs = [None] * (N - 1)        # R_i (accurate) or L_i (fast)
    prev = None
    for i in range(N - 1):
        if fused:
            prev = facs[i] = K.chi2_env(prev, B[i])        # L_i in one launch
            continue
        l, _, r = B[i].shape
        if accurate:
            M = B[i].reshape(l * 2, r) if prev is None else K.gemm(prev, B[i].reshape(l, 2 * r)).reshape(-1, r)
            _, prev = K.qr(M, want_q=False)
        else:
            Bm = B[i].reshape(l * 2, r)
            X = Bm if prev is None else K.gemm(prev, B[i].reshape(l, 2 * r)).reshape(l * 2, r)
            prev = K.gemm(Bm, X, transA=True)              # L_i = B^H (L_{i-1} (x) 1) B
        facs[i] = prev
    # right->left truncation with padded bond 2
    C = K.zeros((N, 8))
    bond = K.zeros((max(N - 1, 1),), dtype=_i32(K))
    ambiguous = K.zeros((1,), dtype=_i32(K))
    b_last = B[N - 1].shape[0]
    T = K.zeros((b_last, 4))
    K.scale_copy(B[N - 1].reshape(b_last * 2, 1), out=T.reshape(b_last * 2, 2)[:, 0:1])
    S4 = K.zeros((4,), dtype=_f64(K))
    Vh4 = K.zeros((4, 4))
    Vsel = K.zeros((4, 2))
    for i in range(N - 1, 0, -1):
        if fused:
            T = K.chi2_bond(facs[i - 1], T, B[i - 1], C[i], bond[i - 1:i], ambiguous)
            continue
        M = K.gemm(facs[i - 1], T)                         # R T (k x 4)  or  L T (b x 4)
        if accurate:
            if M.shape[0] < 4:
                S4.zero_()
                Vh4.zero_()
            K.svd(M, want_u=False, out_s=S4, out_vh=Vh4)
            K.chi2_select(S4, Vh4, C[i], Vsel, bond[i - 1:i])
        else:
            H = K.gemm(T, M, transA=True)                  # T^H L T, 4 x 4 Hermitian PSD, diagonalised in the select kernel
            K.chi2_select(S4, H, C[i], Vsel, bond[i - 1:i], squared=2, ambiguous=ambiguous)
        W = K.gemm(T, Vsel)                                # (b x 2)
        l0, _, b = B[i - 1].shape
        T = K.gemm(B[i - 1].reshape(l0 * 2, b), W).reshape(l0, 4)
    K.chi2_first(T, C[0])
    if not accurate and K.read_int(ambiguous, expect=0):
        return chi2_layer(K, B, debug, accurate=True)
    gates, kinds, bad = K.complete_unitaries(C, bond, N)
    kinds_h = K.read_kinds(kinds, N)
    if K.read_int(bad, expect=0):
        raise ValueError("All the generated unitaries must be unitary.")     # mps.py:838-839
    if debug is not None:
        debug["C"] = C
        debug["bond"] = bond
        debug["accurate"] = accurate
    return gates, kinds_h


def _i32(K):
    import torch
    return torch.int32


def _f64(K):
    import torch
    return torch.float64


def blocks_from_kinds(kinds):
    """[(start, end)] : a block ends at the site carrying the one-qubit gate."""
    out = []
    s = 0
    for i, k in enumerate(kinds):
        if k == 1:
            out.append((s, i))
            s = i + 1
    return out


# --------------------------------------------------------------------------------------
# A6  inverse layer application       mps.py:933-971 (quimb gate_ / gate_split_)
# --------------------------------------------------------------------------------------
def apply_inverse_layer(K, B, gates, kinds, spectra=None, inverse=True, split="svd"):
    """In place on the list ``B``.  theta = G^H (A_i A_{i+1}); SVD, cutoff 1e-10 'rsum2'
    with Frobenius renorm, sqrt(s) to both sides, no max_bond.  ``inverse=False`` is the
    forward application (mps.py:893-931): G itself, sites in ascending order.

    ``split="exact"`` (opt-in, not the reference's arithmetic): the reference never truncates
    here (cutoff only), and everything downstream -- the chi=2 truncation through left
    environments, the |0..0> overlap, the next contraction -- depends on the STATE, not on the
    gauge of the two site tensors.  So theta is re-split trivially, theta = theta . 1 or 1 . theta
    with bond min(2l, 2r) (the rank a generic state has anyway): no SVD, identical gates.  Not
    available: gate-split spectra, and the 1e-10 rank cut for low-rank (structured) states, whose
    bonds then grow to min(2l, 2r)."""
    for s, e in blocks_from_kinds(kinds):
        order = range(e, s - 1, -1) if inverse else range(s, e + 1)
        for i in order:
            G = gates[i]
            if i == e:
                l, _, r = B[i].shape
                K.site_gate(B[i], l, r, G, dagger=inverse)
            else:
                l, _, b = B[i].shape
                _, _, r = B[i + 1].shape
                fused = split == "svd" and _fused_small(K, [B[i], B[i + 1]])
                if fused:
                    X = K.theta_small(B[i], B[i + 1], G, dagger=inverse)
                else:
                    X = K.gemm(B[i].reshape(l * 2, b), B[i + 1].reshape(b, 2 * r))
                    K.theta_gate(X, l, r, G, dagger=inverse)
                if split == "exact":
                    if 2 * l <= 2 * r:
                        B[i] = K.eye(2 * l).reshape(l, 2, 2 * l)
                        B[i + 1] = X.reshape(2 * l, 2, r)
                    else:
                        B[i] = X.reshape(l, 2, 2 * r)
                        B[i + 1] = K.eye(2 * r).reshape(2 * r, 2, r)
                    continue
                U, S, Vh = K.svd(X, backmult=True)
                k = S.shape[0]
                if spectra is not None:
                    spectra.append(S)
                if fused:
                    left, right = K.split_absorb(U, S, Vh, CUTOFF, MODE_RSUM2, 0, k)
                    B[i] = left.reshape(l, 2, k)
                    B[i + 1] = right.reshape(k, 2, r)
                    continue
                rank, f = K.trim(S, k, CUTOFF, MODE_RSUM2)
                n = K.read_int(rank, expect=k)
                B[i] = K.scale_copy(U[:, :n], S, f, mode=2, half_power=True).reshape(l, 2, n)
                B[i + 1] = K.scale_copy(Vh[:n], S, f, mode=1, half_power=True).reshape(n, 2, r)
    return B


# --------------------------------------------------------------------------------------
# A7  overlap with |0..0>             mps.py:1020-1039
# --------------------------------------------------------------------------------------
def zero_overlap(K, B, break_tol=None):
    """conj(psi[0]) as a product of the p=0 slices (only element 0 of the dense vector is used).
    With ``break_tol`` the value only feeds the early-break test (sequential.py:390); in the
    kernels' static mode that test is validated on the device and None is returned."""
    if _fused_small(K, B):
        K.zero_overlap_fused(B, break_tol if break_tol is not None else -1.0)
        return None
    v = B[0][:, 0, :]
    for i in range(1, len(B)):
        v = K.gemm(v, B[i][:, 0, :])
    z = K.read_overlap(v, break_tol if break_tol is not None else -1.0)
    return None if z is None else complex(np.conj(z))


# --------------------------------------------------------------------------------------
# A8/A9  dense optimisation sweeps    sequential.py:400-541
# --------------------------------------------------------------------------------------
def flat_schedule(kinds_per_layer, n_sites):
    sites, kinds = [], []
    for kl in kinds_per_layer:
        sites.extend(range(n_sites))
        kinds.extend(kl)
    return sites, kinds


STORED_SWEEP_MAX_BYTES = 64 << 30      # intermediates kept in HBM up to 64 GiB (180 GB per B200)


def optimize_layers(K, target, gates_all, kinds_per_layer, n_sites, num_sweeps, envs=None, stored=True, small=True,
                    persistent=True):
    """``_optimize_unitary_layers``: per sweep rebuild the dense circuit state from the
    current gates (sequential.py:533, 443-447; the reference's full-rank re-compression
    of that state into an MPS is an identity and is skipped) and run one environment
    sweep (sequential.py:452-505).  ``gates_all``: [L*N, 16] in application order."""
    sites, kinds = flat_schedule(kinds_per_layer, n_sites)
    if (small and num_sweeps > 0 and n_sites <= K.SMALL_SWEEP_MAX_SITES and len(sites) <= K.SMALL_SWEEP_MAX_GATES):
        # both dense vectors fit one SM's shared memory: every sweep in a single launch
        K.sweeps_small(target, n_sites, gates_all, sites, kinds, num_sweeps, 1, envs)
        return gates_all
    stored_bytes = (len(sites) + 1) * 16 * (1 << n_sites)
    use_stored = stored and stored_bytes <= STORED_SWEEP_MAX_BYTES
    if (persistent and use_stored and num_sweeps > 0 and not getattr(K, "static", False)
            and hasattr(K, "sweeps_persist")):
        # every sweep (forward + backward pass) in one persistent cooperative launch, one grid barrier per gate-step
        if K.sweeps_persist(target, n_sites, gates_all, sites, kinds, num_sweeps, envs):
            return gates_all
    c = None
    vwarm = K.zeros((len(sites), 16)) if use_stored else None      # warm start of the 4x4 polar, per gate
    for _ in range(num_sweeps):
        tbar = K.conj_scale_copy(target, conj=True)
        if use_stored:
            # every intermediate c_k stays in HBM; one fused pass per gate in the sweep
            c = K.circuit_states(n_sites, gates_all, sites, kinds, out=c)
            K.sweep_stored(c, tbar, n_sites, gates_all, sites, kinds, envs, vwarm)
        else:
            c = K.circuit_state(n_sites, gates_all, sites, kinds, out=c)
            K.sweep(c, tbar, n_sites, gates_all, sites, kinds, envs)
    return gates_all


# --------------------------------------------------------------------------------------
# A0  top level                       base.py:96-104, sequential.py:330-398, 543-600
# --------------------------------------------------------------------------------------
def from_dense_truncated(K, psi, n_sites, chi, spectra=None):
    """A1 + A2 in one right-to-left pass: the TT-SVD is run in Schmidt form (U.S carried to the
    left, V^H kept) with the 'rel' 1e-10 cutoff and max_bond=chi applied at every split.  Each SVD
    is then the Schmidt decomposition of the state whose right part is already truncated, which is
    exactly what ``compress_right(from_dense(psi), chi)`` computes bond by bond, so the result is the
    same right-canonical MPS (norm on site 0) up to bond phases -- without the QR sweep and the
    second round of SVDs.  (The reference's intermediate 'rsum2' cut inside from_dense only removes
    weight <= 1e-10 and is not reproduced here; ``build_mps(fused=False)`` is the two-pass form.)"""
    N = int(n_sites)
    A = [None] * N
    T = psi.reshape(-1, 1)
    r = 1
    for i in range(N - 1, 0, -1):
        U, S, Vh = K.svd(T.reshape(2 ** i, 2 * r), backmult=True)
        k = S.shape[0]
        if spectra is not None:
            spectra.append(S)
        if getattr(K, "static", False) and getattr(K, "fused_small", True) and hasattr(K, "split_absorb"):
            n = min(k, chi) if chi else k
            T, right = K.split_absorb(U, S, Vh, CUTOFF, MODE_REL, chi or 0, n)
            A[i] = right.reshape(n, 2, r)
            r = n
            continue
        rank, _ = K.trim(S, k, CUTOFF, MODE_REL, chi or 0)
        n = K.read_int(rank, expect=min(k, chi) if chi else k)
        A[i] = K.scale_copy(Vh[:n]).reshape(n, 2, r)
        T = K.scale_copy(U[:, :n], S, None, mode=2, half_power=False)
        r = n
    A[0] = T.reshape(1, 2, r)
    return A


def build_mps(K, psi, n_sites, chi, record=None, fused=True):
    """``MPS.from_statevector`` (mps.py:218-249).  Returns the chi-truncated
    right-canonical MPS (norm on site 0, not renormalised).  ``fused=False`` follows the
    reference's two steps literally (exact TT-SVD with sqrt(s) on both sides, then compression)."""
    rec = record if record is not None else {}
    if fused:
        return from_dense_truncated(K, psi, n_sites, chi, rec.setdefault("truncate", []))
    A = from_dense(K, psi, n_sites, rec.setdefault("tt_svd", []))
    return canonicalize_truncate(K, A, chi, rec.setdefault("truncate", []))


def disentangle(K, A, num_layers, threshold, record=None, split="svd", preconditioned=True):
    """``_get_unitary_layers`` (sequential.py:330-398).  Returns (gates_all [L*N,16] in
    application order, kinds_per_layer, overlaps).

    ``preconditioned``: ``A`` is right-canonical with the 1e-10 'rel' cutoff already applied to every
    bond (what :func:`build_mps` / ``MPS.compress`` return).  Only then is the reference's
    pre-conditioning (sequential.py:360-376: copy, normalize, compress(mode="right"), permute,
    canonicalize("right", normalize=True)) the same tensors with site 0 divided by its norm (the gauge
    is unique up to bond phases).  For any other input -- left-canonical, mixed after
    ``apply_unitary_layer``, built from raw arrays -- the QR sweep + cutoff-only SVD sweep are run:
    they return the right-canonical form with the STATE norm on site 0, which is then normalised."""
    import torch
    rec = record if record is not None else {}
    N = len(A)
    if preconditioned:
        B = copy_mps(K, A)
    else:
        B = canonicalize_truncate(K, A)                               # mps.py:451: left_canonize + right_compress
    normalize_site0(K, B)                                             # mps.py:302-310, :398
    layer_gates, layer_kinds, overlaps = [], [], []
    for _ in range(num_layers):
        gates, kinds = chi2_layer(K, B)                               # mps.py:849-891
        sp = []
        apply_inverse_layer(K, B, gates, kinds, sp, split=split)      # sequential.py:326
        rec.setdefault("gate_split", []).append(sp)
        layer_gates.append(gates)
        layer_kinds.append(kinds)
        # np.isclose(f, 1+0j, atol=1-threshold): |f - 1| <= atol + rtol*|1| with numpy's rtol = 1e-5
        f = zero_overlap(K, B, break_tol=(1 - threshold) + 1e-5)
        overlaps.append(f)
        if f is not None and np.isclose(f, 1 + 0j, atol=1 - threshold):   # sequential.py:390
            break
    layer_gates.reverse()                                             # sequential.py:396
    layer_kinds.reverse()
    gates_all = torch.cat(layer_gates, dim=0).contiguous()
    return gates_all, layer_kinds, overlaps


SCHEDULES = ("DallOall", "IterDiOall", "IterDiOi")


def disentangle_iterative(K, A, num_layers, num_sweeps, threshold, schedule, record=None, split="svd",
                          preconditioned=True):
    """The two schedules the reference names as future work (notebook :459; docstring sequential.py:410,
    428-432; Rudolph et al. 2022), composed from the same stages as the default one:

    ``IterDiOi``    generate layer k from the residual MPS (A4/A5), optimise it ALONE for ``num_sweeps`` sweeps --
                    the rest of the circuit is fixed, so its environments are those of the one-layer circuit
                    against the dense residual -- then take the optimised layer out of the residual (A6).
    ``IterDiOall``  generate layer k from the residual, optimise ALL layers so far against the target (A8/A9),
                    rebuild the residual from the pre-conditioned MPS with the optimised circuit (k x A6).

    Returns what :func:`disentangle` returns, sweeps already done.  With ``num_sweeps = 0`` both are the default
    schedule, launch for launch."""
    import torch
    if schedule not in SCHEDULES[1:]:
        raise ValueError("`schedule` must be one of %s." % (SCHEDULES,))
    rec = record if record is not None else {}
    N = len(A)
    B0 = copy_mps(K, A) if preconditioned else canonicalize_truncate(K, A)
    normalize_site0(K, B0)
    rebuild = schedule == "IterDiOall" and num_sweeps > 0
    B = copy_mps(K, B0) if rebuild else B0
    target = to_dense(K, A) if rebuild else None                     # sequential.py:440 (mps.mps, not normalised)
    layer_gates, layer_kinds, overlaps = [], [], []                   # application order: newest layer first
    for _ in range(num_layers):
        gates, kinds = chi2_layer(K, B)
        sp = []
        if schedule == "IterDiOi":
            if num_sweeps > 0:
                gates = gates.contiguous()
                optimize_layers(K, to_dense(K, B), gates, [kinds], N, num_sweeps)
            layer_gates.insert(0, gates)
            layer_kinds.insert(0, kinds)
            apply_inverse_layer(K, B, gates, kinds, sp, split=split)
        else:
            layer_gates.insert(0, gates)
            layer_kinds.insert(0, kinds)
            if rebuild:
                gates_all = torch.cat(layer_gates, dim=0).contiguous()
                optimize_layers(K, target, gates_all, layer_kinds, N, num_sweeps)
                layer_gates = [gates_all[j * N:(j + 1) * N] for j in range(len(layer_kinds))]
                B = copy_mps(K, B0)
                for g, kd in zip(reversed(layer_gates), reversed(layer_kinds)):   # the layer applied last comes off first
                    sp = []
                    apply_inverse_layer(K, B, g, kd, sp, split=split)
            else:
                apply_inverse_layer(K, B, gates, kinds, sp, split=split)
        rec.setdefault("gate_split", []).append(sp)
        f = zero_overlap(K, B, break_tol=(1 - threshold) + 1e-5)
        overlaps.append(f)
        if f is not None and np.isclose(f, 1 + 0j, atol=1 - threshold):
            break
    gates_all = torch.cat(layer_gates, dim=0).contiguous()
    return gates_all, layer_kinds, overlaps


def prepare_layers_device(K, psi, n_sites, chi, num_layers, threshold=1 - 1e-6, record=None, fused=True, split="svd"):
    """First half of :func:`prepare_device`: normalise ``psi`` in place, build the MPS and extract the layers.
    Returns (gates_all, kinds per layer, overlaps, MPS).  graphs.py captures this much per state and leaves the
    sweeps of the whole batch to one launch."""
    K.div_sqrt(psi, K.vdot(psi, psi))                                 # quick Ket normalisation
    A = build_mps(K, psi, int(n_sites), chi, record, fused)
    gates_all, layer_kinds, overlaps = disentangle(K, A, num_layers, threshold, record, split)
    return gates_all, layer_kinds, overlaps, A


def to_dense_batch(K, A):
    """:func:`to_dense` for W states: A[i] (W, l, 2, r) -> (W, 2^N)."""
    W = A[0].shape[0]
    x = A[0].reshape(W, -1, A[0].shape[3])
    for i in range(1, len(A)):
        _, l, _, r = A[i].shape
        x = K.gemm_batch(x, A[i].reshape(W, l, 2 * r)).reshape(W, -1, r)
    return x.reshape(W, -1)


def chi2_layer_lockstep(K, B, flags):
    """Fast path of :func:`chi2_layer` for W states in ONE launch per step (static mode; B[i]: (W, l, 2, r) with all
    bonds <= FUSED_MAX_BOND).  Same kernels, same arithmetic per state; an ambiguous truncation, a block structure
    other than one block spanning all sites or a failed unitarity check sets the state's flag (the state is then redone
    eagerly, where the accurate path and the error live).  Returns (gates (W, N, 16), kinds of a layer)."""
    import torch
    N = len(B)
    W = B[0].shape[0]
    facs = [None] * (N - 1)
    prev = None
    for i in range(N - 1):
        prev = facs[i] = K.chi2_env_batch(prev, B[i])
    C = K.zeros((W, N, 8))
    bond = K.zeros((W, max(N - 1, 1)), dtype=torch.int32)
    ambiguous = K.zeros((W,), dtype=torch.int32)
    b_last = B[N - 1].shape[1]
    T = K.zeros((W, b_last, 4))
    T.view(W, b_last * 2, 2)[:, :, 0].copy_(B[N - 1].reshape(W, b_last * 2))
    for i in range(N - 1, 0, -1):
        T = K.chi2_bond_batch(facs[i - 1], T, B[i - 1], C, i, bond, i - 1, ambiguous)
    K.chi2_first_batch(T, C)
    K.expect_ints_batch(ambiguous.view(W, 1), 1, 0, flags)
    gates, kinds, bad = K.complete_unitaries_batch(C, bond, N)
    if N > 1:
        K.expect_ints_batch(kinds, N - 1, 2, flags)
    K.expect_ints_batch(kinds[:, N - 1:], 1, 1, flags)
    K.expect_ints_batch(bad.view(W, 1), 1, 0, flags)
    return gates, [2] * (N - 1) + [1]


def prepare_layers_lockstep(K, psis, n_sites, chi, num_layers, threshold, flags):
    """:func:`prepare_layers_device` for W small states AT ONCE in the static (CUDA-graph) mode of one stream.
    A graph lane executes its kernels one after the other, the kernels of a small register are single CTAs, and the
    front end retires a bounded number of graph nodes per second: one state per lane keeps one SM busy and a batch
    is bound by its node count.  Here the W states advance in lock step -- same shapes by the static assumptions --
    and EVERY step is one launch for all of them (grid = W or grid.y = W; the *_batch entries of the C ABI): the
    single-CTA Jacobi SVDs, the split + absorb, the chi=2 truncation, the contraction + gate, the overlap.
    ``psis``: (W, 2^N) device tensor, normalised in place.  ``flags``: int32[W + 1], flags[w] = state w's "an
    assumption failed" flag, flags[W] the group's (a batched SVD did not converge).
    Returns (gates (W, L*N, 16) in application order, [kinds] * L, A) with A[i] the batched MPS sites (W, l, 2, r)."""
    import torch
    assert K.static and chi
    W, N = int(psis.shape[0]), int(n_sites)
    grp = flags[W:W + 1]
    fl = flags[:W]
    for w in range(W):
        K.mismatch = flags[w:w + 1]
        K.div_sqrt(psis[w], K.vdot(psis[w], psis[w]))                 # quick Ket normalisation
    # ---- A1 + A2: TT-SVD in Schmidt form with max_bond (from_dense_truncated) ----
    A = [None] * N
    r = 1
    X = psis.reshape(W, 2 ** (N - 1), 2)
    Tl = None
    for i in range(N - 1, 0, -1):
        m, nn = 2 ** i, 2 * r
        k = min(m, nn)
        K.mismatch = grp
        U, S, Vh = K.svd_small_batch(X, backmult=True)
        n = min(k, chi)
        Xn = K.empty((W, m // 2, 2 * n)) if i > 1 else None
        Tl, right = K.split_absorb_batch(U, S, Vh, CUTOFF, MODE_REL, chi, n, fl, out_left=Xn)
        A[i] = right.reshape(W, n, 2, r)
        X, r = Xn, n
    A[0] = Tl.reshape(W, 1, 2, r)
    # ---- layers (disentangle) ----
    B = [a.clone() for a in A]
    for w in range(W):
        K.mismatch = flags[w:w + 1]
        a0 = B[0][w]
        K.div_sqrt(a0, K.vdot(a0, a0))                                # normalize_site0
    layer_gates = []
    kinds = None
    for _ in range(num_layers):
        G, kinds = chi2_layer_lockstep(K, B, fl)                      # mps.py:849-891
        layer_gates.append(G)
        # inverse layer (mps.py:944-971), one block spanning all sites by the static assumption
        K.site_gate_batch(B[N - 1], G[:, N - 1], dagger=True)
        for i in range(N - 2, -1, -1):
            l, rr = B[i].shape[1], B[i + 1].shape[3]
            Xt = K.theta_small_batch(B[i], B[i + 1], G[:, i], True)
            K.mismatch = grp
            U, S, Vh = K.svd_small_batch(Xt, backmult=True)
            k = S.shape[1]
            left, right = K.split_absorb_batch(U, S, Vh, CUTOFF, MODE_RSUM2, 0, k, fl)
            B[i] = left.reshape(W, l, 2, k)
            B[i + 1] = right.reshape(W, k, 2, rr)
        K.zero_overlap_batch(B, (1 - threshold) + 1e-5, fl)           # sequential.py:390 validated on the device
    gates_all = torch.cat(list(reversed(layer_gates)), dim=1).contiguous()        # sequential.py:396
    return gates_all, [kinds] * num_layers, A


def prepare_device(K, psi, n_sites, chi, num_layers, num_sweeps, threshold=1 - 1e-6, record=None, fused=True,
                   split="svd"):
    """Device-to-device core of :func:`prepare`: ``psi`` is a device vector (overwritten by its
    normalised copy), returns device tensors (gates_all [L*N,16], kinds per layer, overlap [2] =
    <psi|circuit> as (re, im), overlaps list).  No host transfer; this is what graphs.py captures."""
    N = int(n_sites)
    gates_all, layer_kinds, overlaps, A = prepare_layers_device(K, psi, N, chi, num_layers, threshold, record, fused,
                                                                split)
    if num_sweeps > 0:
        target = to_dense(K, A)                                       # sequential.py:440 (mps.mps)
        optimize_layers(K, target, gates_all, layer_kinds, N, num_sweeps)
    sites, kinds = flat_schedule(layer_kinds, N)
    c = K.circuit_state(N, gates_all, sites, kinds)
    return gates_all, layer_kinds, K.vdot(psi, c), overlaps, A


def prepare(K, psi_host, n_sites, chi, num_layers=1, num_sweeps=0, threshold=1 - 1e-6, record=None,
            mps=None, fused=True, split="svd", mps_preconditioned=False, schedule="DallOall"):
    """Whole path on the device.  ``psi_host``: complex128 numpy vector or device tensor.
    Returns dict(gates [L,N,16] numpy, kinds [L][N], n_layers, overlaps, fidelity).
    ``mps``: start from these site tensors instead of building them (``prepare_mps``); any gauge unless
    ``mps_preconditioned`` says they come from :func:`build_mps`."""
    N = int(n_sites)
    if hasattr(psi_host, "data_ptr"):                 # already a device tensor (bench: inputs resident in HBM)
        psi = K.scale_copy(psi_host.reshape(-1, 1)).reshape(-1)
    else:
        psi = K.from_host(np.asarray(psi_host, dtype=np.complex128).reshape(-1))
    if schedule not in SCHEDULES:
        raise ValueError("`schedule` must be one of %s." % (SCHEDULES,))
    if mps is None and schedule == "DallOall":
        gates_all, layer_kinds, ovt, overlaps, A = prepare_device(K, psi, N, chi, num_layers, num_sweeps,
                                                                  threshold, record, fused, split)
    else:
        if mps is None:
            K.div_sqrt(psi, K.vdot(psi, psi))                         # quick Ket normalisation
            A = build_mps(K, psi, N, chi, record, fused)
            mps_preconditioned = True
        else:
            A = mps
        if schedule == "DallOall":
            gates_all, layer_kinds, overlaps = disentangle(K, A, num_layers, threshold, record, split,
                                                           preconditioned=mps_preconditioned)
            if num_sweeps > 0:
                optimize_layers(K, to_dense(K, A), gates_all, layer_kinds, N, num_sweeps)
        else:
            gates_all, layer_kinds, overlaps = disentangle_iterative(K, A, num_layers, num_sweeps, threshold,
                                                                     schedule, record, split, mps_preconditioned)
        sites, kinds = flat_schedule(layer_kinds, N)
        ovt = K.vdot(psi, K.circuit_state(N, gates_all, sites, kinds))
    ov = K.to_host(ovt)
    if hasattr(K, "check_small_svd"):
        K.check_small_svd()
    L = len(layer_kinds)
    return {
        "gates": K.to_host(gates_all).reshape(L, N, 16),
        "kinds": layer_kinds,
        "n_layers": L,
        "overlaps": overlaps,
        "fidelity": float(np.hypot(ov[0], ov[1])),
        "overlap": (float(ov[0]), float(ov[1])),
        "n_sites": N,
        "mps": A,
    }
