"""U3 + CX circuit backend: depth and op counts in the basis the reference reports them in.

The reference hands every extracted 4x4 / 2x2 unitary to ``circuit.unitary(matrix, qubits)``
(qmprs/synthesis/mps_encoding/sequential.py:182-187) and quick's ``QiskitCircuit`` lowers it to
U3 and CX gates; README.md:68-70 and the notebook (``State Preparation using MPS
Sequential.ipynb`` cells 21-28) quote ``get_depth()`` / ``count_ops()`` of that lowering
(10 q, 15 layers: depth 223, 405 CX = 3 per two-qubit gate, 1095 U3 = 8 per two-qubit gate + 1 per
one-qubit gate).  quick is not installable here, so this module restates the published
constructions from scratch to make "same gate count and depth"
checkable without it:

* one-qubit gate  -> one U3 (ZYZ Euler angles) and a global phase;
* two-qubit gate  -> Cartan (KAK) decomposition in the magic basis, interaction vector moved into
  the Weyl chamber, then the fewest CX its class needs: generic -> the three-CX circuit of Vatan &
  Williams (2004) for exp(i(a XX + b YY + c ZZ)), 3 CX + 8 U3; c = 0 -> 2 CX + 6 U3; CX class
  -> 1 CX + 4 U3; tensor product -> 2 U3 (the notebook's 4-qubit line, 123 U3 / 44 CX for 15
  two-qubit gates, shows quick also drops a CX and two U3 for a special gate).

``U3CXCircuit.get_depth()`` merges runs of one-qubit gates on a wire (one level per run), which is
the rule that reproduces every depth the reference publishes for generic states
(6 N + 12 L - 17: 73 @ 5 q/5 layers, 115 @ 6 q/8, 133 @ 9 q/8, 223 @ 10 q/15, 235 @ 12 q/15);
``count_ops()`` counts the emitted gates unmerged, as quick does.  Host-side numpy, outside the
accelerated path (the accelerated path ends at ``circuit.unitary``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["u3_matrix", "u3_angles", "kak_decompose", "two_qubit_ops", "U3CXCircuit"]

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_I = np.eye(2, dtype=np.complex128)
# columns = magic (Bell) basis; conjugation by it maps SU(2) x SU(2) onto SO(4)
_MAGIC = np.array([[1, 0, 0, 1j], [0, 1j, 1, 0], [0, 1j, -1, 0], [1, 0, 0, -1j]], dtype=np.complex128) / np.sqrt(2)
_XX, _YY, _ZZ = np.kron(_X, _X), np.kron(_Y, _Y), np.kron(_Z, _Z)
# a XX + b YY + c ZZ is diagonal in the magic basis: phases = _WEYL @ (g, a, b, c)
_WEYL = np.stack([np.ones(4)] + [np.real(np.diag(_MAGIC.conj().T @ P @ _MAGIC)) for P in (_XX, _YY, _ZZ)], axis=1)
# kron(A, B): first factor = more significant index bit.  CX with control on the first / second factor:
_CX_FIRST = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
_CX_SECOND = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)


def u3_matrix(theta: float, phi: float, lam: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -np.exp(1j * lam) * s],
                     [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]], dtype=np.complex128)


def u3_angles(u) -> tuple[float, float, float, float]:
    """(theta, phi, lam, phase) with u = exp(i phase) * U3(theta, phi, lam)."""
    u = np.asarray(u, dtype=np.complex128)
    c, s = abs(u[0, 0]), abs(u[1, 0])
    theta = 2.0 * np.arctan2(s, c)
    if c >= s:
        phase = np.angle(u[0, 0])
        sum_pl = np.angle(u[1, 1]) - phase                      # phi + lam
        phi = np.angle(u[1, 0]) - phase if s > 1e-14 else 0.0
        lam = sum_pl - phi
    else:
        phi_p = np.angle(u[1, 0])                               # phi + phase
        lam_p = np.angle(-u[0, 1])                              # lam + phase
        phase = np.angle(u[0, 0]) if c > 1e-14 else 0.0
        phi, lam = phi_p - phase, lam_p - phase
    return float(theta), float(phi), float(lam), float(phase)


def _rz(t):
    return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])


def _ry(t):
    return np.array([[np.cos(t / 2), -np.sin(t / 2)], [np.sin(t / 2), np.cos(t / 2)]], dtype=np.complex128)


def _split_kron(k):
    """k = phase * kron(a, b) with a, b in SU(2); returns (a, b, phase)."""
    t = k.reshape(2, 2, 2, 2)                                   # [i, k, j, l] = a[i, j] b[k, l]
    i, kk, j, l = np.unravel_index(np.argmax(np.abs(t)), t.shape)
    a = t[:, kk, :, l].copy()
    b = t[i, :, j, :].copy()
    a /= np.sqrt(np.linalg.det(a))
    b /= np.sqrt(np.linalg.det(b))
    ab = np.kron(a, b)
    idx = np.argmax(np.abs(ab))
    phase = k.flat[idx] / ab.flat[idx]
    return a, b, phase


def kak_decompose(u, rng=None):
    """u = phase * kron(a1, a2) @ exp(i (a XX + b YY + c ZZ)) @ kron(b1, b2).

    Returns ``(phase, a1, a2, (a, b, c), b1, b2)``; a1, a2, b1, b2 in SU(2).  Construction: in the
    magic basis u' = u / det(u)^(1/4) becomes O1 D O2 with O1, O2 in SO(4) and D diagonal; O2 comes
    from the simultaneous real diagonalisation of Re and Im of the symmetric unitary u_m^T u_m.
    """
    u = np.asarray(u, dtype=np.complex128)
    if u.shape != (4, 4) or not np.allclose(u @ u.conj().T, np.eye(4), atol=1e-8):
        raise ValueError("kak_decompose needs a 4x4 unitary")
    rng = rng or np.random.default_rng(12345)
    ph0 = np.linalg.det(u) ** 0.25
    um = _MAGIC.conj().T @ (u / ph0) @ _MAGIC
    m2 = um.T @ um
    p = None
    for attempt in range(64):
        r = 0.5 * np.pi * (attempt / 64.0) if attempt < 2 else rng.uniform(0, np.pi)
        h = np.cos(r) * m2.real + np.sin(r) * m2.imag
        _, cand = np.linalg.eigh(0.5 * (h + h.T))
        d2 = cand.T @ m2 @ cand
        if np.abs(d2 - np.diag(np.diag(d2))).max() < 1e-9:
            p = cand
            break
    if p is None:
        raise np.linalg.LinAlgError("simultaneous diagonalisation failed")
    if np.linalg.det(p) < 0:
        p[:, 0] = -p[:, 0]
    half = 0.5 * np.angle(np.diag(p.T @ m2 @ p))
    if np.cos(half.sum()) < 0:                                   # det(D) must be +1
        half[0] += np.pi
    o1 = um @ p @ np.diag(np.exp(-1j * half))
    if np.abs(o1.imag).max() > 1e-7:
        raise np.linalg.LinAlgError("left factor is not real orthogonal")
    o1 = o1.real
    g, a, b, c = np.linalg.solve(_WEYL, half)
    a1, a2, pl = _split_kron(_MAGIC @ o1 @ _MAGIC.conj().T)
    b1, b2, pr = _split_kron(_MAGIC @ p.T @ _MAGIC.conj().T)
    phase = ph0 * np.exp(1j * g) * pl * pr
    return phase, a1, a2, (float(a), float(b), float(c)), b1, b2


def _weyl_core(a, b, c):
    """Three-CX circuit equal to exp(i(a XX + b YY + c ZZ)) up to a global phase, as a list of layers
    in time order; each layer is ("1q", first_factor_gate, second_factor_gate) or ("cx", control)."""
    return [
        ("1q", _I, _rz(-0.5 * np.pi)),
        ("cx", 1),
        ("1q", _I, _ry(2 * a + 0.5 * np.pi)),
        ("cx", 0),
        ("1q", _rz(-2 * c - 0.5 * np.pi), _ry(-2 * b - 0.5 * np.pi)),
        ("cx", 1),
        ("1q", _rz(0.5 * np.pi), _I),
    ]


def _layers_matrix(layers):
    m = np.eye(4, dtype=np.complex128)
    for lay in layers:
        if lay[0] == "1q":
            m = np.kron(lay[1], lay[2]) @ m
        else:
            m = (_CX_FIRST if lay[1] == 0 else _CX_SECOND) @ m
    return m


def _canonicalize(v, A, B, phase):
    """Move the interaction vector v = (a, b, c) of phase * A @ N(v) @ B into the Weyl chamber
    pi/4 >= a >= b >= |c| by the symmetries of N, updating the 4x4 local factors A and B."""
    v = np.array(v, dtype=float)
    paulis = (_X, _Y, _Z)
    for k in range(3):                                   # N(v) = N(v - s pi/2 e_k) (i P_k x P_k)^s
        s = int(np.round(v[k] / (0.5 * np.pi)))
        if s:
            v[k] -= s * 0.5 * np.pi
            pk = np.linalg.matrix_power(paulis[k], s % 2)
            B = np.kron(pk, pk) @ B
            phase = phase * (1j ** s)

    def swap(k, l, A, B):                                # (V x V) exchanges P_k x P_k and P_l x P_l
        vv = (paulis[k] + paulis[l]) / np.sqrt(2)
        v[[k, l]] = v[[l, k]]
        return A @ np.kron(vv, vv), np.kron(vv, vv) @ B

    def negate(k, l, A, B):                              # (P_m x 1) flips the signs of v_k and v_l
        pm = np.kron(paulis[3 - k - l], _I)
        v[k], v[l] = -v[k], -v[l]
        return A @ pm, pm @ B

    for i in range(3):                                   # sort by decreasing |v|
        j = i + int(np.argmax(np.abs(v[i:])))
        if j != i:
            A, B = swap(i, j, A, B)
    if v[0] < 0 and v[1] < 0:
        A, B = negate(0, 1, A, B)
    elif v[0] < 0:
        A, B = negate(0, 2, A, B)
    elif v[1] < 0:
        A, B = negate(1, 2, A, B)
    return v, A, B, phase


_CX_KAK = None


def _cx_in_chamber():
    """CX (control on the first factor) = phase * A @ N(pi/4, 0, 0) @ B, computed once."""
    global _CX_KAK
    if _CX_KAK is None:
        ph, a1, a2, v, b1, b2 = kak_decompose(_CX_FIRST)
        v, A, B, ph = _canonicalize(v, np.kron(a1, a2), np.kron(b1, b2), ph)
        assert np.allclose(v, [0.25 * np.pi, 0, 0], atol=1e-9)
        _CX_KAK = (ph, A, B)
    return _CX_KAK


def _emit(layers, u, ops_out):
    """Append the gates of `layers` and return the global phase that makes them equal to u."""
    full = _layers_matrix(layers)
    idx = np.argmax(np.abs(full))
    gphase = u.flat[idx] / full.flat[idx]
    if not np.allclose(gphase * full, u, atol=1e-7):
        raise np.linalg.LinAlgError("two-qubit decomposition failed to reproduce the gate")
    for lay in layers:
        if lay[0] == "cx":
            ops_out.append(("CX", (lay[1], 1 - lay[1])))
        else:
            for w, g in ((0, lay[1]), (1, lay[2])):
                th, ph, la, gp = u3_angles(g)
                gphase = gphase * np.exp(1j * gp)
                ops_out.append(("U3", (th, ph, la), w))
    return float(np.angle(gphase))


def two_qubit_ops(u, atol=1e-9):
    """Lower a 4x4 unitary to [("U3", (theta, phi, lam), w) | ("CX", (control_w, target_w))] on the
    local wires w in {0, 1} (wire 0 = first kron factor = more significant matrix-index bit) plus a
    global phase, with the fewest CX its Weyl-chamber coordinates (a, b, c) allow:
    (0,0,0) -> 2 U3; (pi/4,0,0) -> 1 CX + 4 U3; c = 0 -> 2 CX + 6 U3; generic -> 3 CX + 8 U3."""
    u = np.asarray(u, dtype=np.complex128)
    ops = []
    t = u.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(4, 4)        # operator-Schmidt matrix
    sv = np.linalg.svd(t, compute_uv=False)
    if sv[1] <= 1e-12 * sv[0]:
        a, b, _ = _split_kron(u)
        return ops, _emit([("1q", a, b)], u, ops)
    phase, a1, a2, v, b1, b2 = kak_decompose(u)
    v, A, B, phase = _canonicalize(v, np.kron(a1, a2), np.kron(b1, b2), phase)
    (a1, a2, _), (b1, b2, _) = _split_kron(A), _split_kron(B)
    if abs(v[2]) > atol:                                                 # generic: three CX
        core = _weyl_core(*v)
        core[0] = ("1q", core[0][1] @ b1, core[0][2] @ b2)
        core[-1] = ("1q", a1 @ core[-1][1], a2 @ core[-1][2])
    elif abs(v[0] - 0.25 * np.pi) <= atol and abs(v[1]) <= atol:         # CX class: one CX
        _, Ac, Bc = _cx_in_chamber()
        (l1, l2, _), (r1, r2, _) = _split_kron(A @ Ac.conj().T), _split_kron(Bc.conj().T @ B)
        core = [("1q", r1, r2), ("cx", 0), ("1q", l1, l2)]
    else:                                                                # c = 0: two CX
        # N(a, b, 0) = (V x V) N(a, 0, b) (V x V), V = (Y + Z)/sqrt(2);  N(a, 0, b) = CX (Rx(-2a) x Rz(-2b)) CX
        vv = (_Y + _Z) / np.sqrt(2)
        rx = np.array([[np.cos(v[0]), 1j * np.sin(v[0])], [1j * np.sin(v[0]), np.cos(v[0])]])
        core = [("1q", vv @ b1, vv @ b2), ("cx", 0), ("1q", rx, _rz(-2 * v[1])), ("cx", 0),
                ("1q", a1 @ vv, a2 @ vv)]
    return ops, _emit(core, u, ops)


class U3CXCircuit:
    """Duck-typed ``circuit_framework`` (sequential.py:182-187, 208) that lowers every unitary to
    U3 / CX on arrival.  Qubit 0 is the least significant bit of the statevector index and the
    first listed qubit of a two-qubit gate is the less significant bit of its 4x4 matrix index
    (the convention of :class:`qmprs_b200.circuit.GateListCircuit`)."""

    def __init__(self, num_qubits: int) -> None:
        self.num_qubits = int(num_qubits)
        self.ops: list[tuple] = []          # ("U3", (theta, phi, lam), q) | ("CX", (control, target))
        self.global_phase = 0.0
        self._n_unitary = {1: 0, 2: 0}
        self._n_phase = 0

    def unitary(self, matrix, qubit_indices) -> None:
        q = [int(qubit_indices)] if np.isscalar(qubit_indices) else [int(x) for x in qubit_indices]
        m = np.array(matrix, dtype=np.complex128)
        if m.shape != (2 ** len(q), 2 ** len(q)) or len(q) not in (1, 2):
            raise ValueError("matrix shape does not match the number of qubits")
        if len(q) == 1:
            th, ph, la, gp = u3_angles(m)
            self.ops.append(("U3", (th, ph, la), q[0]))
        else:
            ops, gp = two_qubit_ops(m)
            wire = {0: q[1], 1: q[0]}       # first kron factor = more significant bit = second listed qubit
            for op in ops:
                if op[0] == "U3":
                    self.ops.append(("U3", op[1], wire[op[2]]))
                else:
                    self.ops.append(("CX", (wire[op[1][0]], wire[op[1][1]])))
        self._n_unitary[len(q)] += 1
        if abs(gp) > 1e-15:
            self._n_phase += 1
        self.global_phase += gp

    def count_ops(self) -> dict:
        out = {"U3": 0, "CX": 0}
        for op in self.ops:
            out[op[0]] += 1
        out["GlobalPhase"] = self._n_phase
        out["unitary1"], out["unitary2"] = self._n_unitary[1], self._n_unitary[2]
        return out

    def get_depth(self) -> int:
        """Depth in the U3/CX basis with runs of one-qubit gates on a wire merged into one level."""
        level = [0] * self.num_qubits
        open_1q = [False] * self.num_qubits       # last gate on the wire is a one-qubit gate
        for op in self.ops:
            if op[0] == "U3":
                w = op[2]
                if not open_1q[w]:
                    level[w] += 1
                    open_1q[w] = True
            else:
                c, t = op[1]
                d = max(level[c], level[t]) + 1
                level[c] = level[t] = d
                open_1q[c] = open_1q[t] = False
        return max(level) if level else 0

    def get_statevector(self) -> np.ndarray:
        n = self.num_qubits
        psi = np.zeros([2] * n, dtype=np.complex128)
        psi[(0,) * n] = 1.0
        for op in self.ops:
            if op[0] == "U3":
                ax = n - 1 - op[2]
                psi = np.moveaxis(np.tensordot(u3_matrix(*op[1]), psi, axes=([1], [ax])), 0, ax)
            else:
                axc, axt = n - 1 - op[1][0], n - 1 - op[1][1]
                sl = [slice(None)] * n
                sl[axc] = 1
                sub = psi[tuple(sl)]
                psi = psi.copy()
                psi[tuple(sl)] = np.flip(sub, axis=axt - (1 if axt > axc else 0))
        return np.exp(1j * self.global_phase) * np.ascontiguousarray(psi).reshape(-1)
