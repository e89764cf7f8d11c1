"""Thin Python binding of the C ABI (include/qmprs_b200.h) on torch CUDA tensors.

PyTorch is used for device memory and streams only; every arithmetic step is one of
the hand-written sm_100a kernels in ``csrc/``.  There is no CPU path: constructing
:class:`CudaKernels` without CUDA or without the built library raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

C128 = torch.complex128
F64 = torch.float64
I32 = torch.int32

CUTOFF = 1e-10
TIE_REL = 1e-6
SIGN_TOL = 1e-12
MODE_REL, MODE_RSUM2 = 0, 1
CHI2_AMBIGUOUS_REL = 1e-3   # squared chi=2 path: s_1 <= 1e-3 s_0 -> redo the layer with the QR-based path


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_process_device = None      # the library keeps per-process state (opt-in smem sizes, side streams, split-K workspace)


class CudaKernels:
    """Kernel handle.  One DEVICE per process (one process per GPU under torchrun): the library's
    function attributes, side streams and split-K workspaces are created once per process on the first
    device used, so a handle for a second device in the same process is refused instead of failing at
    launch time.  Methods take/return torch tensors on that device."""

    def __init__(self, device="cuda:0"):
        global _process_device
        if not torch.cuda.is_available():
            raise RuntimeError("qmprs_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback.")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if _process_device is None:
            _process_device = self.device
        elif _process_device != self.device:
            raise RuntimeError(f"qmprs_b200 drives one GPU per process: this process already uses {_process_device}, "
                               f"refusing {self.device} (launch one process per GPU, e.g. torchrun)")
        torch.cuda.set_device(self.device)
        self._svd_work = None
        self._sweep_work = None
        self.launches = 0          # C-ABI calls issued (each launches >= 1 kernel)
        self.svd_sweeps = 0
        self.svd_log = None        # set to a list to record (m, n, sweeps, backmult) of every multi-CTA SVD (scripts/svd_census.py)
        self.svd_unconverged = 0
        self.svd_tol = 1e-14
        self.svd_max_sweeps = 30
        # matrices that fit one SM's shared memory are decomposed by the single-CTA solver (qm_svd_small): one
        # launch, no host synchronisation per sweep.  Its sticky "did not converge" flag is read by check_small_svd().
        self.small_svd = True
        self._small_flag = None
        # speculative static-shape mode (graphs.py): no host read-backs; assumptions are validated on
        # the device and recorded in `mismatch`
        self.static = False
        self.static_sweeps = 12
        self.static_sweeps_small = 6
        self.mismatch = None

    # ---- memory / plumbing --------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, shape, dtype=C128):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=C128):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def eye(self, n, dtype=C128):
        return torch.eye(n, dtype=dtype, device=self.device)

    def from_host(self, arr, dtype=C128):
        return torch.as_tensor(np.ascontiguousarray(arr)).to(dtype).to(self.device)

    def to_host(self, t):
        return t.detach().cpu().numpy()

    def read_int(self, t, expect=None):
        """Scalar that decides a shape.  Static mode returns the assumed value and enqueues the check."""
        if self.static:
            assert expect is not None, "static mode needs the assumed value"
            self._check(self.lib.qm_expect_ints(_p(t), None, 1, int(expect), _p(self.mismatch), self._stream()),
                        "qm_expect_ints")
            return int(expect)
        return int(t.item())

    def read_kinds(self, kinds, n_sites):
        """Per-site gate kinds of a layer; static mode assumes one block spanning all sites."""
        if self.static:
            st = self._stream()
            self._check(self.lib.qm_expect_ints(_p(kinds), None, n_sites - 1, 2, _p(self.mismatch), st),
                        "qm_expect_ints")
            self._check(self.lib.qm_expect_ints(_p(kinds[n_sites - 1:]), None, 1, 1, _p(self.mismatch), st),
                        "qm_expect_ints")
            return [2] * (n_sites - 1) + [1]
        return [int(x) for x in self.to_host(kinds)]

    def read_overlap(self, v, tol):
        """<0..0|psi> as a host complex (eager) or None (static: the early break is assumed not to
        fire, i.e. |f - 1| > tol is validated on the device)."""
        if self.static:
            self._check(self.lib.qm_expect_not_close(_p(v), float(tol), _p(self.mismatch), self._stream()),
                        "qm_expect_not_close")
            return None
        return complex(self.to_host(v).reshape(-1)[0])

    def begin_static(self):
        self.mismatch = self.zeros((1,), I32)
        self.static = True

    def end_static(self):
        self.static = False

    def synchronize(self):
        torch.cuda.synchronize(self.device)

    def _check(self, code, name):
        self.launches += 1
        if code != 0:
            raise RuntimeError(f"{name} failed with status {code}")

    @staticmethod
    def _ld(t):
        assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "matrix views must have unit column stride"
        return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))

    # ---- dense linear algebra -----------------------------------------------------
    def gemm(self, A, B, out=None, transA=False):
        """out = op(A) @ B, op(A) = A or A^H (A stored k x m when transA)."""
        if transA:
            k, m = A.shape
        else:
            m, k = A.shape
        k2, n = B.shape
        assert k == k2
        if out is None:
            out = self.empty((m, n))
        self._check(self.lib.qm_zgemm(m, n, k, 1.0, 0.0, _p(A), self._ld(A), _p(B), self._ld(B), 0.0, 0.0,
                                      _p(out), self._ld(out), 1, 0, 0, 0, 1 if transA else 0, self._stream()),
                    "qm_zgemm")
        return out

    # tall / wide matrices: Gram pre-conditioning (see _svd_preconditioned)
    PRECOND_ASPECT = 8
    PRECOND_MIN = 64

    def _svd_preconditioned(self, A):
        """SVD of a tall A (m >= 8n) in three GEMM-rich steps around two much cheaper Jacobi runs:
        G = A^H A = V L V^H by the Jacobi SVD of the small n x n matrix; A1 = A V has nearly orthogonal
        columns, so the full-accuracy one-sided Jacobi on A1 (which is what fixes the accuracy the Gram
        matrix squared away) needs 2-3 sweeps instead of 10-13 over the long vectors; Vh = V1^H V^H."""
        G = self.gemm(A, A, transA=True)                  # n x n Hermitian PSD
        # V from the ACCUMULATED rotations (Vh of the Jacobi run: unitary to rounding whatever the spectrum);
        # the normalised-row factor U loses orthogonality for eigenvalues near eps * |G|.
        # (a pre-conditioner: eigenvalues below eps |G| never meet the relative tolerance; bounded, no warning)
        _, _, Vhg = self.svd(G, _plain=True, _max_sweeps=12)
        A1 = self.gemm(A, self.transpose(Vhg, conj=True))
        U, S, Vh1 = self.svd(A1, _plain=True)
        return U, S, self.gemm(Vh1, Vhg)

    def svd(self, A, want_u=True, want_vh=True, out_s=None, out_vh=None, _plain=False, backmult=False,
            _max_sweeps=None):
        """Thin SVD.  ``backmult`` (the MPS path's splits): QM_SVD_BACKMULT of the C ABI -- the rotations are not
        accumulated, the second factor comes from one ZGEMM against ``A`` (include/qmprs_b200.h)."""
        m, n = A.shape
        k = min(m, n)
        flags = 1 if backmult else 0
        if (not _plain and want_u and want_vh and out_s is None and out_vh is None and k >= self.PRECOND_MIN
                and max(m, n) >= self.PRECOND_ASPECT * k):
            if m >= n:                         # (accumulating mode inside: its Gram stage needs a unitary V whatever the spectrum)
                return self._svd_preconditioned(A)
            # wide: A^H = V S U^H
            V, S, Uh = self._svd_preconditioned(self.transpose(A, conj=True))
            return self.transpose(Uh, conj=True), S, self.transpose(V, conj=True)
        if self.small_svd and self.lib.qm_svd_small_fits(m, n, flags):
            U = self.empty((m, k)) if want_u else None
            S = out_s if out_s is not None else self.empty((k,), F64)
            Vh = out_vh if out_vh is not None else (self.empty((k, n)) if want_vh else None)
            if self.static:
                flag = self.mismatch
            else:
                if self._small_flag is None:
                    self._small_flag = self.zeros((1,), I32)
                flag = self._small_flag
            self._check(self.lib.qm_svd_small(m, n, _p(A), self._ld(A), 0, _p(U), k, 0, _p(S), 0, _p(Vh),
                                              (self._ld(Vh) if Vh is not None else n), 0, self.svd_tol,
                                              self.svd_max_sweeps, flags, 1, _p(flag), self._stream()), "qm_svd_small")
            return U, S, Vh
        need = int(self.lib.qm_svd_work_bytes(m, n))
        if self._svd_work is None or self._svd_work.numel() < need:
            self._svd_work = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=self.device)
        U = self.empty((m, k)) if want_u else None
        S = out_s if out_s is not None else self.empty((k,), F64)
        Vh = out_vh if out_vh is not None else (self.empty((k, n)) if want_vh else None)
        if self.static:
            # single-block problems (<= 32 short vectors) converge in 3-4 Gram iterations
            fixed = self.static_sweeps if k > 32 else self.static_sweeps_small
            self._check(self.lib.qm_svd_static(m, n, _p(A), self._ld(A), _p(U), k, _p(S), _p(Vh),
                                               (self._ld(Vh) if Vh is not None else n), _p(self._svd_work),
                                               self._svd_work.numel(), self.svd_tol, fixed,
                                               _p(self.mismatch), flags, self._stream()), "qm_svd_static")
            return U, S, Vh
        info = (ctypes.c_int * 2)()
        self._check(self.lib.qm_svd(m, n, _p(A), self._ld(A), _p(U), k, _p(S), _p(Vh),
                                    (self._ld(Vh) if Vh is not None else n), _p(self._svd_work),
                                    self._svd_work.numel(), self.svd_tol, _max_sweeps or self.svd_max_sweeps, info,
                                    flags, self._stream()), "qm_svd")
        self.svd_sweeps += info[0]
        if self.svd_log is not None:
            self.svd_log.append((m, n, int(info[0]), bool(backmult)))
        if not info[1] and _max_sweeps is None:
            # a partially converged decomposition would silently drive rank cut-offs and gates
            import warnings
            self.svd_unconverged += 1
            warnings.warn(f"qm_svd({m}x{n}) did not converge to tol={self.svd_tol} in {self.svd_max_sweeps} sweeps",
                          RuntimeWarning, stacklevel=2)
        return U, S, Vh

    def svd_small_batch(self, X, backmult=False):
        """Thin SVDs of the W same-shape matrices X[w] (W, m, n) in ONE launch, one CTA each (qm_svd_small).
        Returns U (W, m, k), S (W, k), Vh (W, k, n); non-convergence of any of them sets ``self.mismatch``."""
        W, m, n = X.shape
        k = min(m, n)
        flags = 1 if backmult else 0
        assert X.is_contiguous() and self.lib.qm_svd_small_fits(m, n, flags)
        U = self.empty((W, m, k))
        S = self.empty((W, k), F64)
        Vh = self.empty((W, k, n))
        flag = self.mismatch if self.static else self._small_flag_tensor()
        self._check(self.lib.qm_svd_small(m, n, _p(X), n, m * n, _p(U), k, m * k, _p(S), k, _p(Vh), n, k * n, self.svd_tol,
                                          self.svd_max_sweeps, flags, W, _p(flag), self._stream()), "qm_svd_small")
        return U, S, Vh

    def _small_flag_tensor(self):
        if self._small_flag is None:
            self._small_flag = self.zeros((1,), I32)
        return self._small_flag

    def check_small_svd(self):
        """Read (and clear) the sticky non-convergence flag of the single-CTA SVDs issued since the last check."""
        if self._small_flag is None:
            return
        if int(self._small_flag.item()):
            import warnings
            self.svd_unconverged += 1
            self._small_flag.zero_()
            warnings.warn(f"a qm_svd_small problem did not converge to tol={self.svd_tol} in {self.svd_max_sweeps} sweeps",
                          RuntimeWarning, stacklevel=2)

    def transpose(self, A, conj=False):
        rows, cols = A.shape
        out = self.empty((cols, rows))
        self._check(self.lib.qm_transpose(_p(out), rows, _p(A), self._ld(A), rows, cols, 1 if conj else 0,
                                          self._stream()), "qm_transpose")
        return out

    QR_BLOCKED_MIN = 64       # below: the column-by-column kernels (a panel is 32 columns)

    def qr(self, A, want_q=True, blocked=None):
        """Reduced QR with non-negative diag(R).  Returns (Q or None, R).  Matrices with more than 64 columns
        are factored in panels of 32 with the trailing updates as ZGEMMs on the tensor cores (qm_qr_blocked)."""
        m, n = A.shape
        k = min(m, n)
        if blocked is None:
            blocked = k >= self.QR_BLOCKED_MIN
        F = self.empty((m, n))
        self._check(self.lib.qm_scale_copy(_p(F), n, _p(A), self._ld(A), m, n, None, None, 0, 0, self._stream()),
                    "qm_scale_copy")
        tau = self.empty((k,))
        work = None
        if blocked:
            need = int(self.lib.qm_qr_work_bytes(m, n))
            work = self.__dict__.get("_qr_work")
            if work is None or work.numel() < need:
                work = self._qr_work = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._check(self.lib.qm_qr_blocked(m, n, _p(F), n, _p(tau), _p(work), work.numel(), self._stream()),
                        "qm_qr_blocked")
        else:
            self._check(self.lib.qm_qr(m, n, _p(F), n, _p(tau), self._stream()), "qm_qr")
        Q = None
        if want_q:
            Q = self.empty((m, k))
            if blocked:
                self._check(self.lib.qm_qr_formq_blocked(m, k, _p(F), n, _p(tau), _p(Q), k, _p(work), work.numel(),
                                                         self._stream()), "qm_qr_formq_blocked")
            else:
                self._check(self.lib.qm_qr_formq(m, k, _p(F), n, _p(tau), _p(Q), k, self._stream()), "qm_qr_formq")
        R = self.empty((k, n))
        self._check(self.lib.qm_qr_finish(m, n, _p(F), n, _p(R), n, _p(Q), k, self._stream()), "qm_qr_finish")
        return Q, R

    # ---- MPS bookkeeping ----------------------------------------------------------
    def trim(self, S, k, cutoff, mode, max_bond=0):
        rank = self.empty((1,), I32)
        f = self.empty((1,), F64)
        self._check(self.lib.qm_trim(_p(S), k, cutoff, mode, int(max_bond or 0), _p(rank), _p(f), self._stream()),
                    "qm_trim")
        return rank, f

    def scale_copy(self, inp, S=None, f=None, mode=0, half_power=False, out=None):
        rows, cols = inp.shape
        if out is None:
            out = self.empty((rows, cols))
        self._check(self.lib.qm_scale_copy(_p(out), self._ld(out), _p(inp), self._ld(inp), rows, cols, _p(S), _p(f),
                                           mode, 1 if half_power else 0, self._stream()), "qm_scale_copy")
        return out

    def theta_gate(self, X, l, r, G, dagger):
        self._check(self.lib.qm_theta_gate(_p(X), l, r, _p(G), 1 if dagger else 0, self._stream()), "qm_theta_gate")

    def site_gate(self, B, l, r, G, dagger):
        self._check(self.lib.qm_site_gate(_p(B), l, r, _p(G), 1 if dagger else 0, self._stream()), "qm_site_gate")

    def chi2_select(self, S4, Vh4, Csite, Vsel, bond_slot, squared=False, ambiguous=None):
        self._check(self.lib.qm_chi2_select(_p(S4), _p(Vh4), 4, CUTOFF, TIE_REL, _p(Csite), _p(Vsel), _p(bond_slot),
                                            int(squared), CHI2_AMBIGUOUS_REL, _p(ambiguous), self._stream()),
                    "qm_chi2_select")

    def chi2_first(self, T0, Csite):
        self._check(self.lib.qm_chi2_first(_p(T0), _p(Csite), self._stream()), "qm_chi2_first")

    def complete_unitaries(self, C, bond, n_sites):
        gates = self.empty((n_sites, 16))
        kinds = self.empty((n_sites,), I32)
        bad = self.zeros((1,), I32)
        self._check(self.lib.qm_complete_unitaries(_p(C), _p(bond), n_sites, _p(gates), _p(kinds), _p(bad), SIGN_TOL,
                                                   self._stream()), "qm_complete_unitaries")
        return gates, kinds, bad

    # ---- fused bookkeeping for small registers (csrc/small_mps.cu) ---------------------
    FUSED_MAX_BOND = 64

    def split_absorb(self, U, S, Vh, cutoff, mode, max_bond, expect, out_left=None):
        """trim + rank check + absorb in one launch (static mode): returns (left m x expect, right expect x n).
        ``out_left``: contiguous buffer of m * expect elements that receives ``left`` (e.g. the next split's input)."""
        m, k = U.shape
        n = Vh.shape[1]
        left = self.empty((m, expect)) if out_left is None else out_left.reshape(m, expect)
        right = self.empty((expect, n))
        self._check(self.lib.qm_split_absorb(_p(U), self._ld(U), _p(S), _p(Vh), self._ld(Vh), m, n, k, float(cutoff),
                                             int(mode), int(max_bond or 0), int(expect), _p(left), _p(right),
                                             _p(self.mismatch), self._stream()), "qm_split_absorb")
        return left, right

    def theta_small(self, A, A2, G, dagger, out=None):
        l, _, b = A.shape
        r = A2.shape[2]
        X = self.empty((2 * l, 2 * r)) if out is None else out
        self._check(self.lib.qm_theta_small(_p(A), _p(A2), l, b, r, _p(G), 1 if dagger else 0, _p(X), self._stream()),
                    "qm_theta_small")
        return X

    def chi2_env(self, Lprev, B):
        l, _, r = B.shape
        out = self.empty((r, r))
        self._check(self.lib.qm_chi2_env(_p(Lprev), _p(B), l, r, _p(out), self._stream()), "qm_chi2_env")
        return out

    def chi2_bond(self, L, T, Bprev, Csite, bond_slot, ambiguous):
        b = T.shape[0]
        l0 = Bprev.shape[0]
        Tout = self.empty((l0, 4))
        self._check(self.lib.qm_chi2_bond(_p(L), b, _p(T), _p(Bprev), l0, CUTOFF, TIE_REL, CHI2_AMBIGUOUS_REL, _p(Csite),
                                          _p(bond_slot), _p(ambiguous), _p(Tout), self._stream()), "qm_chi2_bond")
        return Tout

    def zero_overlap_fused(self, B, tol):
        """<0..0|psi> in one launch; with tol >= 0 (static mode) the early break is validated on the device."""
        n = len(B)
        ptrs = (ctypes.c_void_p * n)(*[b.data_ptr() for b in B])
        dims = (ctypes.c_int * (n + 1))(*([int(B[0].shape[0])] + [int(b.shape[2]) for b in B]))
        out = self.empty((1,))
        self._check(self.lib.qm_zero_overlap(ptrs, dims, n, float(tol), _p(out), _p(self.mismatch), self._stream()),
                    "qm_zero_overlap")
        return out

    # ---- the same for W same-shape states in one launch (lock-step lanes of graphs.py) --------------------
    @staticmethod
    def _strides(*vals):
        return (ctypes.c_longlong * len(vals))(*[int(v) for v in vals])

    @staticmethod
    def _bstride(t):
        """Element stride between the states of a batched tensor whose per-state block is dense."""
        assert t[0].is_contiguous(), "per-state blocks of a batched tensor must be dense"
        return t.stride(0) if t.shape[0] > 1 else t[0].numel()

    def gemm_batch(self, A, B):
        """out[w] = A[w] @ B[w] (W, m, k) x (W, k, n), one launch (qm_zgemm's batch argument)."""
        W, m, k = A.shape
        n = B.shape[2]
        assert B.shape[0] == W and B.shape[1] == k
        out = self.empty((W, m, n))
        self._check(self.lib.qm_zgemm(m, n, k, 1.0, 0.0, _p(A), k, _p(B), n, 0.0, 0.0, _p(out), n, W,
                                      self._bstride(A), self._bstride(B), m * n, 0, self._stream()), "qm_zgemm")
        return out

    def split_absorb_batch(self, U, S, Vh, cutoff, mode, max_bond, expect, flags, out_left=None):
        """:meth:`split_absorb` on U (W, m, k), S (W, k), Vh (W, k, n): left (W, m, expect), right (W, expect, n);
        ``flags``: int32[W], one "assumption failed" flag per state."""
        W, m, k = U.shape
        n = Vh.shape[2]
        left = self.empty((W, m, expect)) if out_left is None else out_left
        right = self.empty((W, expect, n))
        assert left.is_contiguous() and left.numel() == W * m * expect
        self._check(self.lib.qm_split_absorb_batch(
            _p(U), k, _p(S), _p(Vh), n, m, n, k, float(cutoff), int(mode), int(max_bond or 0), int(expect), _p(left),
            _p(right), _p(flags), W, self._strides(self._bstride(U), self._bstride(S), self._bstride(Vh), m * expect,
                                                   expect * n), self._stream()), "qm_split_absorb_batch")
        return left.reshape(W, m, expect), right

    def theta_small_batch(self, A, A2, G, dagger):
        """A (W, l, 2, b), A2 (W, b, 2, r), G (W, 16) view (any state stride) -> theta (W, 2l, 2r)."""
        W, l, _, b = A.shape
        r = A2.shape[3]
        X = self.empty((W, 2 * l, 2 * r))
        self._check(self.lib.qm_theta_small_batch(
            _p(A), _p(A2), l, b, r, _p(G), 1 if dagger else 0, _p(X), W,
            self._strides(self._bstride(A), self._bstride(A2), G.stride(0), 4 * l * r), self._stream()),
            "qm_theta_small_batch")
        return X

    def chi2_env_batch(self, Lprev, B):
        W, l, _, r = B.shape
        out = self.empty((W, r, r))
        self._check(self.lib.qm_chi2_env_batch(
            _p(Lprev), _p(B), l, r, _p(out), W,
            self._strides(0 if Lprev is None else self._bstride(Lprev), self._bstride(B), r * r), self._stream()),
            "qm_chi2_env_batch")
        return out

    def chi2_bond_batch(self, L, T, Bprev, C, site, bond, slot, ambiguous):
        """One bond of the chi=2 truncation for W states: C (W, N, 8) receives site ``site``, bond (W, nb) slot
        ``slot``, ambiguous int32[W]."""
        W, b, _ = T.shape
        l0 = Bprev.shape[1]
        Tout = self.empty((W, l0, 4))
        Cs, bs = C[:, site], bond[:, slot]
        self._check(self.lib.qm_chi2_bond_batch(
            _p(L), b, _p(T), _p(Bprev), l0, CUTOFF, TIE_REL, CHI2_AMBIGUOUS_REL, _p(Cs), _p(bs), _p(ambiguous), _p(Tout),
            W, self._strides(self._bstride(L), self._bstride(T), self._bstride(Bprev), C.stride(0), bond.stride(0), 1,
                             l0 * 4), self._stream()), "qm_chi2_bond_batch")
        return Tout

    def chi2_first_batch(self, T, C):
        W = T.shape[0]
        self._check(self.lib.qm_chi2_first_batch(_p(T), _p(C), W, self._bstride(T), C.stride(0), self._stream()),
                    "qm_chi2_first_batch")

    def complete_unitaries_batch(self, C, bond, n_sites):
        W = C.shape[0]
        assert C.is_contiguous() and C.shape[1] == n_sites
        gates = self.empty((W, n_sites, 16))
        kinds = self.empty((W, n_sites), I32)
        bad = self.zeros((W,), I32)
        self._check(self.lib.qm_complete_unitaries_batch(_p(C), _p(bond), n_sites, _p(gates), _p(kinds), _p(bad), SIGN_TOL,
                                                         W, bond.stride(0), self._stream()),
                    "qm_complete_unitaries_batch")
        return gates, kinds, bad

    def expect_ints_batch(self, vals, n, scalar, flags):
        """vals: (W, >= n) int32 view; flags[w] = 1 where one of the first n values of row w is not ``scalar``."""
        W = vals.shape[0]
        self._check(self.lib.qm_expect_ints_batch(_p(vals), None, int(n), int(scalar), _p(flags), W, vals.stride(0),
                                                  self._stream()), "qm_expect_ints_batch")

    def site_gate_batch(self, B, G, dagger):
        """B (W, l, 2, r) in place, G (W, 16) view (first four entries = the 2x2 gate)."""
        W, l, _, r = B.shape
        self._check(self.lib.qm_site_gate_batch(_p(B), l, r, _p(G), 1 if dagger else 0, W, self._bstride(B), G.stride(0),
                                                self._stream()), "qm_site_gate_batch")

    def zero_overlap_batch(self, B, tol, flags):
        """<0..0|psi_w> for the W states of the batched sites B[i] (W, l, 2, r); early-break check per state."""
        n = len(B)
        W = B[0].shape[0]
        ptrs = (ctypes.c_void_p * n)(*[b.data_ptr() for b in B])
        dims = (ctypes.c_int * (n + 1))(*([int(B[0].shape[1])] + [int(b.shape[3]) for b in B]))
        out = self.empty((W,))
        self._check(self.lib.qm_zero_overlap_batch(ptrs, dims, n, float(tol), _p(out), _p(flags), W,
                                                   self._strides(*[self._bstride(b) for b in B]), self._stream()),
                    "qm_zero_overlap_batch")
        return out

    def reverse3(self, a):
        l, _, r = a.shape
        out = self.empty((r, 2, l))
        self._check(self.lib.qm_reverse3(_p(out), _p(a), l, r, self._stream()), "qm_reverse3")
        return out

    # ---- vectors --------------------------------------------------------------------
    def conj_scale_copy(self, inp, conj=False, scale=1.0):
        out = self.empty(inp.shape)
        self._check(self.lib.qm_conj_scale_copy(_p(out), _p(inp), inp.numel(), 1 if conj else 0, float(scale),
                                                self._stream()), "qm_conj_scale_copy")
        return out

    def vdot(self, a, b):
        """(re, im) of sum conj(a) b; reproducible (per-CTA partials live behind the result, fixed-order sum)."""
        out = self.empty((int(self.lib.qm_vdot_out_doubles()),), F64)
        self._check(self.lib.qm_vdot(_p(a), _p(b), a.numel(), _p(out), self._stream()), "qm_vdot")
        return out[:2]

    def div_sqrt(self, x, nrm2):
        self._check(self.lib.qm_div_sqrt(_p(x), x.numel(), _p(nrm2), self._stream()), "qm_div_sqrt")

    # ---- dense statevector path -----------------------------------------------------
    def apply_gate(self, x, n_sites, site, kind, G, op=0):
        self._check(self.lib.qm_apply_gate(_p(x), n_sites, site, kind, _p(G), op, self._stream()), "qm_apply_gate")

    @staticmethod
    def _int_array(v):
        arr = (ctypes.c_int * len(v))(*[int(x) for x in v])
        return arr

    def circuit_state(self, n_sites, gates, sites, kinds, out=None):
        c = out if out is not None else self.empty((1 << n_sites,))
        self._check(self.lib.qm_circuit_state(_p(c), n_sites, _p(gates), self._int_array(sites),
                                              self._int_array(kinds), len(sites), self._stream()), "qm_circuit_state")
        self.launches += len(sites)
        return c

    def sweep(self, c, tbar, n_sites, gates, sites, kinds, envs=None):
        if self._sweep_work is None:
            self._sweep_work = torch.empty(int(self.lib.qm_sweep_work_bytes()), dtype=torch.uint8, device=self.device)
        self._check(self.lib.qm_sweep(_p(c), _p(tbar), n_sites, _p(gates), self._int_array(sites),
                                      self._int_array(kinds), len(sites), _p(self._sweep_work), _p(envs),
                                      self._stream()), "qm_sweep")
        self.launches += 3 * len(sites)


    def circuit_states(self, n_sites, gates, sites, kinds, out=None):
        """All intermediates c_0..c_M ((M+1) x 2^N) of the circuit applied to |0..0>."""
        M = len(sites)
        cs = out if out is not None else self.empty((M + 1, 1 << n_sites))
        self._check(self.lib.qm_circuit_states(_p(cs), n_sites, _p(gates), self._int_array(sites),
                                               self._int_array(kinds), M, self._stream()), "qm_circuit_states")
        return cs

    def sweep_stored(self, cs, tbar, n_sites, gates, sites, kinds, envs=None, vwarm=None):
        if self._sweep_work is None:
            self._sweep_work = torch.empty(int(self.lib.qm_sweep_work_bytes()), dtype=torch.uint8, device=self.device)
        self._check(self.lib.qm_sweep_stored(_p(cs), _p(tbar), n_sites, _p(gates), self._int_array(sites),
                                             self._int_array(kinds), len(sites), _p(self._sweep_work), _p(envs),
                                             _p(vwarm), self._stream()), "qm_sweep_stored")

    def _sched_dev(self, sites, kinds):
        key = (tuple(int(x) for x in sites), tuple(int(x) for x in kinds))
        cache = self.__dict__.setdefault("_sched_cache", {})
        dev = cache.get(key)
        if dev is None:                      # created on first use (the eager warm-up precedes any graph capture)
            dev = (torch.tensor(key[0], dtype=torch.int32, device=self.device),
                   torch.tensor(key[1], dtype=torch.int32, device=self.device))
            cache[key] = dev
        return dev

    def sweeps_persist(self, target, n_sites, gates, sites, kinds, num_sweeps, envs=None):
        """All sweeps of one large register in one persistent cooperative launch (qm_sweeps_persist).
        Returns False when the device cannot co-schedule the grid (caller falls back to the per-gate kernels)."""
        M = len(sites)
        dev = self._sched_dev(sites, kinds)
        need = int(self.lib.qm_sweeps_persist_work_bytes(M))
        w = self.__dict__.get("_persist_work")
        if w is None or w.numel() < need:
            w = self._persist_work = torch.empty(need, dtype=torch.uint8, device=self.device)
        cs = self.empty((M + 1, 1 << n_sites))
        tbar = self.empty((1 << n_sites,))
        code = self.lib.qm_sweeps_persist(_p(cs), _p(tbar), _p(target), n_sites, _p(gates), _p(dev[0]), _p(dev[1]), M,
                                          int(num_sweeps), _p(w), _p(envs), self._stream())
        if code == -3:
            return False
        self._check(code, "qm_sweeps_persist")
        return True

    SMALL_SWEEP_MAX_SITES = 12
    SMALL_SWEEP_MAX_GATES = 256

    def sweeps_small(self, targets, n_sites, gates, sites, kinds, num_sweeps, batch=1, envs=None, psis=None,
                     overlaps=None):
        """All sweeps of `batch` small states in one launch (one CTA per state, vectors in shared memory).
        ``targets``: [batch, 2^N] dense (not conjugated); ``gates``: [batch * M, 16], updated in place;
        ``overlaps`` (optional [batch, 2] float64) receives <psi|circuit|0..0>/|psi| for the final gates, psi from
        ``psis`` ([batch, 2^N]) or the target."""
        dev = self._sched_dev(sites, kinds)
        self._check(self.lib.qm_sweeps_small(_p(targets), n_sites, _p(gates), _p(dev[0]), _p(dev[1]), len(sites),
                                             int(num_sweeps), int(batch), _p(envs), _p(psis), _p(overlaps),
                                             self._stream()), "qm_sweeps_small")

    # ---- instrumentation ------------------------------------------------------------
    def launch_count(self):
        """Kernel launches issued by the library since load (counted in QM_LAUNCH)."""
        return int(self.lib.qm_launch_count())

    def set_pdl(self, on):
        """Programmatic dependent launch on/off; returns the previous setting."""
        return bool(self.lib.qm_set_pdl(1 if on else 0))

    def prof_begin(self):
        self.lib.qm_prof_begin()

    def prof_end(self):
        """{class: (total_ms, launches, algorithmic work)} -- times from CUDA events on the launch stream."""
        n = int(self.lib.qm_prof_num_classes())
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        code = self.lib.qm_prof_end(ms, cnt)
        if code != 0:
            raise RuntimeError(f"qm_prof_end failed with status {code}")
        work = (ctypes.c_double * n)()
        self.lib.qm_prof_work_get(work)
        return {self.lib.qm_prof_class_name(i).decode(): (float(ms[i]), int(cnt[i]), float(work[i]))
                for i in range(n)}


_instances = {}


def get_kernels(device=None):
    """Process-wide kernel handle for ``device`` (default: the current CUDA device, i.e. the
    rank's GPU under torchrun).  Fails loudly without CUDA / the .so."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("qmprs_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback.")
        device = f"cuda:{torch.cuda.current_device()}"
    key = str(device)
    if key not in _instances:
        _instances[key] = CudaKernels(device)
    return _instances[key]
