"""U3/CX lowering (qmprs_b200/transpile.py): the decomposition reproduces the gates, uses the fewest CX
per Weyl class, and reproduces the depths / op counts the reference publishes (README.md:68-70, notebook
cells 27-28) and the depth inequalities of its tests (test_sequential_encoding.py:121, 155)."""
import numpy as np
import pytest
from scipy.linalg import expm
from scipy.stats import unitary_group

from oracle import qmprs_oracle as O
from qmprs_b200.circuit import GateListCircuit
from qmprs_b200.transpile import U3CXCircuit, kak_decompose, two_qubit_ops, u3_angles, u3_matrix
from tests.test_oracle import partial_entanglement_state

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]])
Z = np.diag([1, -1]).astype(complex)
CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
SWAP = np.eye(4)[[0, 2, 1, 3]].astype(complex)


def weyl(a, b, c):
    return expm(1j * (a * np.kron(X, X) + b * np.kron(Y, Y) + c * np.kron(Z, Z)))


def both(n, emit):
    c1, c2 = U3CXCircuit(n), GateListCircuit(n)
    for c in (c1, c2):
        emit(c)
    return c1, c2


def test_u3_angles_roundtrip():
    rng = np.random.default_rng(0)
    cases = [unitary_group.rvs(2, random_state=rng) for _ in range(100)]
    cases += [np.eye(2), X, Y, Z, np.diag([1, 1j]), (X + Z) / np.sqrt(2)]
    for u in cases:
        th, ph, la, gp = u3_angles(u)
        assert np.allclose(np.exp(1j * gp) * u3_matrix(th, ph, la), u, atol=1e-12)


def test_kak_reconstructs():
    rng = np.random.default_rng(1)
    for u in [unitary_group.rvs(4, random_state=rng) for _ in range(50)] + [CX, SWAP, np.eye(4, dtype=complex)]:
        ph, a1, a2, (a, b, c), b1, b2 = kak_decompose(u)
        assert np.allclose(ph * np.kron(a1, a2) @ weyl(a, b, c) @ np.kron(b1, b2), u, atol=1e-9)
        for g in (a1, a2, b1, b2):
            assert abs(np.linalg.det(g) - 1) < 1e-9


def test_two_qubit_lowering_matches_dense_and_uses_fewest_cx():
    rng = np.random.default_rng(2)

    def loc():
        return np.kron(unitary_group.rvs(2, random_state=rng), unitary_group.rvs(2, random_state=rng))

    cases = [(unitary_group.rvs(4, random_state=rng), 3) for _ in range(40)]
    for _ in range(10):
        a, b = rng.uniform(-3, 3, 2)
        cases += [(loc() @ CX @ loc(), 1), (loc() @ np.diag([1, 1, 1, -1]).astype(complex) @ loc(), 1),
                  (loc() @ weyl(a, b, 0) @ loc(), 2), (loc() @ weyl(0, a, b) @ loc(), 2),
                  (loc() @ weyl(a, 0, 0) @ loc(), 2), (loc() @ SWAP @ loc(), 3), (loc(), 0)]
    for u, ncx in cases:
        def emit(c):
            r = np.random.default_rng(7)
            for q in range(3):
                c.unitary(unitary_group.rvs(2, random_state=r), q)
            c.unitary(u, [2, 0])
        c1, c2 = both(3, emit)
        assert np.abs(c1.get_statevector() - c2.get_statevector()).max() < 1e-9
        ops = c1.count_ops()
        assert ops["CX"] == ncx and ops["U3"] == 3 + 2 * (ncx + 1)
        assert len(two_qubit_ops(u)[0]) == 3 * ncx + 2


# notebook cell 27 (config: 5 layers below 6 qubits, 8 below 10, 15 below 14; bond dimension 2^n) prints count_ops()
# per register size and cell 28 the depths [7, 13, 65, 73, 115, 121, 127, 133, 223, 229, 235] for 2..12 qubits: outputs
# of the REAL reference (quimb + quick) on unseeded random states.  For generic states they depend on the structure
# only (layers used, blocks, CX per Weyl class), so every row can be reproduced exactly -- except 4 qubits, where the
# early break and the CX count of the last gates depend on the state (the notebook's own second run, cell 29, has
# depth 53 there instead of 65).
NOTEBOOK_TABLE = [  # n, layers asked, depth, U3, CX
    (2, 5, 7, 9, 3), (3, 5, 13, 17, 6), (5, 5, 73, 165, 60), (6, 8, 115, 328, 120), (7, 8, 121, 392, 144),
    (8, 8, 127, 456, 168), (9, 8, 133, 520, 192), (10, 15, 223, 1095, 405), (11, 15, 229, 1215, 450),
    (12, 15, 235, 1335, 495)]


@pytest.mark.parametrize("n,layers,depth,u3,cx", NOTEBOOK_TABLE)
def test_published_depths_and_op_counts(n, layers, depth, u3, cx):
    # README.md:70 depth 223; notebook cell 27 op counts, cell 28 depths
    psi = O.random_state(n, 0)
    res = O.prepare(psi, n, 2 ** n, layers, 0, gauge="canonical")
    used = len(res["layers"])
    assert used == (layers if n >= 5 else 1)              # 2 and 3 qubits: exact after one layer (sequential.py:390)
    c1, c2 = both(n, lambda c: [c.unitary(g, q) for g, q in O.emit_gates(res["layers"], n)])
    assert c1.get_depth() == depth
    if n >= 5:
        assert depth == 6 * n + 12 * layers - 17
    ops = c1.count_ops()
    n2, n1 = O.count_gates(res["layers"])
    assert (ops["CX"], ops["U3"]) == (3 * n2, 8 * n2 + n1) == (cx, u3)
    assert np.abs(c1.get_statevector() - c2.get_statevector()).max() < 1e-9
    assert abs(abs(np.vdot(psi, c1.get_statevector())) - O.circuit_fidelity(psi, res["layers"], n)) < 1e-9


def test_reference_depth_inequalities():
    # test_sequential_encoding.py:91-121: partial entanglement, 1 layer -> depth <= 20
    psi = partial_entanglement_state()
    res = O.prepare(psi, 8, 32, 1, 0, gauge="canonical")
    c = U3CXCircuit(8)
    for g, q in O.emit_gates(res["layers"], 8):
        c.unitary(g, q)
    assert 1 - abs(np.vdot(psi, c.get_statevector())) < 1e-2
    assert c.get_depth() <= 20
    # test_sequential_encoding.py:123-155: 4 qubits, GHZ-like + H(3), 1 layer, 1 sweep -> depth <= 7
    n = 4
    psi = np.zeros(2 ** n, dtype=complex)
    for b in (0, 1):
        for h in (0, 1):
            psi[(b << 0) | (b << 1) | (b << 2) | (h << 3)] = 1
    psi /= np.linalg.norm(psi)
    res = O.prepare(psi, 4, 16, 1, 1, gauge="canonical")
    c = U3CXCircuit(n)
    for g, q in O.emit_gates(res["layers"], n):
        c.unitary(g, q)
    assert 1 - abs(np.vdot(psi, c.get_statevector())) < 1e-2
    assert c.get_depth() <= 7
