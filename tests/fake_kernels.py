"""TEST DOUBLE for qmprs_b200.kernels.CudaKernels  --  test infrastructure only.

Implements the kernel interface with numpy on CPU torch tensors so that the HOST logic
(qmprs_b200/host.py: launch sequencing, reshapes, index conventions, rank plumbing) can
be exercised in the CPU-only test run (``-m "not gpu"``).  It is never imported by the
package; the product path constructs CudaKernels, which raises without CUDA.
Each method restates the contract of the C-ABI entry point of the same name
(include/qmprs_b200.h), not its implementation.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import qmprs_oracle as O

C128 = torch.complex128
F64 = torch.float64
I32 = torch.int32
CUTOFF = 1e-10
TIE_REL = 1e-6
SIGN_TOL = 1e-12


def _np(t):
    return t.detach().numpy()


class FakeKernels:
    def __init__(self, svd_phase_seed=None):
        self.device = torch.device("cpu")
        self.launches = 0
        self.svd_sweeps = 0
        self._rng = np.random.default_rng(svd_phase_seed) if svd_phase_seed is not None else None

    def empty(self, shape, dtype=C128):
        return torch.zeros(shape, dtype=dtype)

    def zeros(self, shape, dtype=C128):
        return torch.zeros(shape, dtype=dtype)

    def eye(self, n, dtype=C128):
        return torch.eye(n, dtype=dtype)

    def from_host(self, arr, dtype=C128):
        return torch.as_tensor(np.ascontiguousarray(arr)).to(dtype).clone()

    def to_host(self, t):
        return _np(t).copy()

    static = False

    def read_int(self, t, expect=None):
        return int(t.item())

    def read_kinds(self, kinds, n_sites):
        return [int(x) for x in _np(kinds)]

    def read_overlap(self, v, tol):
        return complex(_np(v).reshape(-1)[0])

    def synchronize(self):
        pass

    # ---- linear algebra ----
    def gemm(self, A, B, out=None, transA=False):
        a = np.conj(_np(A)).T if transA else _np(A)
        r = torch.as_tensor(a @ _np(B))
        if out is None:
            return r.contiguous()
        out.copy_(r)
        return out

    def svd(self, A, want_u=True, want_vh=True, out_s=None, out_vh=None, backmult=False):
        u, s, vh = np.linalg.svd(_np(A), full_matrices=False)
        if self._rng is not None:       # arbitrary singular-vector phases, as a Jacobi SVD would return
            ph = np.exp(2j * np.pi * self._rng.random(s.size))
            u = u * ph[None, :]
            vh = vh * np.conj(ph)[:, None]
        k = s.size
        S = out_s if out_s is not None else torch.zeros(k, dtype=F64)
        S[:k] = torch.as_tensor(s)
        Vh = None
        if out_vh is not None:
            out_vh[:k] = torch.as_tensor(vh)
            Vh = out_vh
        elif want_vh:
            Vh = torch.as_tensor(vh).contiguous()
        U = torch.as_tensor(u).contiguous() if want_u else None
        return U, S, Vh

    def qr(self, A, want_q=True):
        q, r = O.qr_pos(_np(A))
        return (torch.as_tensor(q).contiguous() if want_q else None), torch.as_tensor(r).contiguous()

    # ---- bookkeeping ----
    def trim(self, S, k, cutoff, mode, max_bond=0):
        n, f = O.trim(_np(S)[:k], cutoff, "rel" if mode == 0 else "rsum2", max_bond or None)
        return torch.tensor([n], dtype=I32), torch.tensor([f], dtype=F64)

    def scale_copy(self, inp, S=None, f=None, mode=0, half_power=False, out=None):
        x = _np(inp).copy()
        if mode:
            n = x.shape[0] if mode == 1 else x.shape[1]
            s = _np(S)[:n].astype(np.float64).copy()
            if f is not None:
                s = s * float(f.item())
            if half_power:
                s = np.sqrt(s)
            x = x * (s[:, None] if mode == 1 else s[None, :])
        r = torch.as_tensor(x)
        if out is None:
            return r.contiguous()
        out.copy_(r)
        return out

    def theta_gate(self, X, l, r, G, dagger):
        g = _np(G).reshape(4, 4)
        if dagger:
            g = np.conj(g).T
        x = _np(X).reshape(l, 2, 2, r)
        y = np.einsum("abcd,lcdr->labr", g.reshape(2, 2, 2, 2), x)
        X.copy_(torch.as_tensor(y.reshape(2 * l, 2 * r)))

    def site_gate(self, B, l, r, G, dagger):
        g = _np(G).reshape(-1)[:4].reshape(2, 2)
        if dagger:
            g = np.conj(g).T
        y = np.einsum("op,lpr->lor", g, _np(B).reshape(l, 2, r))
        B.copy_(torch.as_tensor(y.reshape(B.shape)))

    def chi2_select(self, S4, Vh4, Csite, Vsel, bond_slot, squared=False, ambiguous=None):
        s = _np(S4)
        if squared == 2:                       # Vh4 is the 4x4 Hermitian matrix: diagonalise it here
            w, v = np.linalg.eigh(_np(Vh4))
            order = np.argsort(w)[::-1]
            s = w[order]
            Vh4 = torch.as_tensor(np.conj(v[:, order]).T.copy())
        if squared:
            s = np.sqrt(np.maximum(s, 0.0))
            if ambiguous is not None and s[1] <= 1e-3 * s[0]:
                ambiguous[0] = 1
        vh = _np(Vh4)
        n = int(np.count_nonzero(s > CUTOFF * s[0]))
        n = min(max(n, 1), 2)
        rows = np.zeros((2, 4), dtype=np.complex128)
        for j in range(n):
            row = vh[j].copy()
            m2 = np.abs(row) ** 2
            pick = int(np.argmax(m2 >= (1 - TIE_REL) * m2.max()))
            a = abs(row[pick])
            ph = row[pick] / a if a > 0 else 1.0
            rows[j] = row / ph
        Csite.copy_(torch.as_tensor(rows.reshape(8)))
        Vsel.copy_(torch.as_tensor(np.conj(rows).T.copy()))
        bond_slot[0] = n

    def chi2_first(self, T0, Csite):
        t = _np(T0).reshape(4)
        out = np.zeros(8, dtype=np.complex128)
        out[:4] = t / np.linalg.norm(t)
        Csite.copy_(torch.as_tensor(out))

    def complete_unitaries(self, C, bond, n_sites):
        c = _np(C).reshape(n_sites, 2, 2, 2)
        b = _np(bond)
        gates = np.zeros((n_sites, 16), dtype=np.complex128)
        kinds = np.zeros(n_sites, dtype=np.int32)
        bad = 0
        for i in range(n_sites):
            dl = 1 if i == 0 else int(b[i - 1])
            dr = 1 if i == n_sites - 1 else int(b[i])
            a = c[i][:dl, :, :dr]
            if dr < 2:
                g = O.last_site_unitary(a, dl < 2, "canonical")
            elif dl < 2:
                g = O.first_site_unitary(a, "canonical")
            else:
                g = O.two_site_unitary(a, "canonical")
            if not O.is_unitary(g):
                bad = 1
            gates[i, : g.size] = g.reshape(-1)
            kinds[i] = 2 if g.shape[0] == 4 else 1
        return torch.as_tensor(gates), torch.as_tensor(kinds), torch.tensor([bad], dtype=I32)

    def reverse3(self, a):
        return torch.as_tensor(np.ascontiguousarray(_np(a).transpose(2, 1, 0)))

    # ---- vectors ----
    def conj_scale_copy(self, inp, conj=False, scale=1.0):
        x = _np(inp)
        return torch.as_tensor((np.conj(x) if conj else x) * scale).contiguous()

    def vdot(self, a, b):
        z = np.vdot(_np(a).reshape(-1), _np(b).reshape(-1))
        return torch.tensor([z.real, z.imag], dtype=F64)

    def div_sqrt(self, x, nrm2):
        x.div_(float(np.sqrt(nrm2[0].item())))

    # ---- dense path ----
    @staticmethod
    def _mat(G, kind, op):
        d = 4 if kind == 2 else 2
        g = _np(G).reshape(-1)[: d * d].reshape(d, d)
        if op == 1:
            g = np.conj(g).T
        elif op == 2:
            g = g.T
        return g

    def apply_gate(self, x, n_sites, site, kind, G, op=0):
        y = O.apply_gate_dense(_np(x), n_sites, site, self._mat(G, kind, op))
        x.copy_(torch.as_tensor(y))

    def circuit_state(self, n_sites, gates, sites, kinds, out=None):
        c = out if out is not None else torch.zeros(1 << n_sites, dtype=C128)
        c.zero_()
        c[0] = 1.0
        for g in range(len(sites)):
            self.apply_gate(c, n_sites, sites[g], kinds[g], gates[g], 0)
        return c

    def circuit_states(self, n_sites, gates, sites, kinds, out=None):
        M = len(sites)
        cs = out if out is not None else torch.zeros((M + 1, 1 << n_sites), dtype=C128)
        cs.zero_()
        cs[0, 0] = 1.0
        for g in range(M):
            cs[g + 1].copy_(cs[g])
            self.apply_gate(cs[g + 1], n_sites, sites[g], kinds[g], gates[g], 0)
        return cs

    def sweep_stored(self, cs, tbar, n_sites, gates, sites, kinds, envs=None, vwarm=None):
        N = n_sites
        for g in range(len(sites) - 1, -1, -1):
            site, kind = sites[g], kinds[g]
            d = 4 if kind == 2 else 2
            k = 2 if kind == 2 else 1
            L, R = 2 ** site, 2 ** (N - site - k)
            E = np.tensordot(_np(tbar).reshape(L, d, R), _np(cs[g]).reshape(L, d, R), axes=([0, 2], [0, 2]))
            Gn = np.conj(O.polar_unitary(E, "canonical"))
            gates[g, : d * d] = torch.as_tensor(Gn.reshape(-1))
            if envs is not None:
                envs[g, : d * d] = torch.as_tensor(E.reshape(-1))
            self.apply_gate(tbar, N, site, kind, gates[g], 2)

    SMALL_SWEEP_MAX_SITES = 12
    SMALL_SWEEP_MAX_GATES = 256

    def sweeps_small(self, targets, n_sites, gates, sites, kinds, num_sweeps, batch=1, envs=None):
        M = len(sites)
        targets = targets.reshape(batch, -1)
        for b in range(batch):
            g = gates[b * M:(b + 1) * M]
            for _ in range(num_sweeps):
                c = self.circuit_state(n_sites, g, sites, kinds)
                tbar = torch.conj(targets[b]).clone()
                self.sweep(c, tbar, n_sites, g, sites, kinds, None if envs is None else envs[b * M:(b + 1) * M])

    def sweep(self, c, tbar, n_sites, gates, sites, kinds, envs=None):
        N = n_sites
        for g in range(len(sites) - 1, -1, -1):
            site, kind = sites[g], kinds[g]
            d = 4 if kind == 2 else 2
            k = 2 if kind == 2 else 1
            self.apply_gate(c, N, site, kind, gates[g], 1)
            L, R = 2 ** site, 2 ** (N - site - k)
            E = np.tensordot(_np(tbar).reshape(L, d, R), _np(c).reshape(L, d, R), axes=([0, 2], [0, 2]))
            Gn = np.conj(O.polar_unitary(E, "canonical"))
            gates[g, : d * d] = torch.as_tensor(Gn.reshape(-1))
            if envs is not None:
                envs[g, : d * d] = torch.as_tensor(E.reshape(-1))
            self.apply_gate(tbar, N, site, kind, gates[g], 2)
