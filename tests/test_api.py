"""Drop-in surface: import paths, signatures and error behaviour of the reference
(qmprs/synthesis/mps_encoding/base.py:58-65, sequential.py:113-153, 588-600;
qmprs/primitives/mps.py:151-216; tests/primitives/test_mps.py:49-88).  No GPU needed:
argument validation happens before any device work, and without CUDA the first
computation must fail loudly instead of falling back to the CPU."""
import numpy as np
import pytest
import torch

from qmprs.primitives import MPS
from qmprs.synthesis.mps_encoding import MPSEncoder, Sequential
from qmprs_b200 import GateListCircuit, Ket
from qmprs_b200.primitives.mps import DeviceMPS


def test_import_paths_and_types():
    import qmprs
    assert issubclass(Sequential, MPSEncoder)
    assert qmprs.primitives.MPS is MPS
    enc = Sequential(GateListCircuit)
    assert enc.circuit_framework is GateListCircuit
    assert enc.fidelity_threshold == 1 - 1e-6


def test_fidelity_threshold_validation():
    enc = Sequential(GateListCircuit)
    enc.fidelity_threshold = 0.5
    assert enc.fidelity_threshold == 0.5
    for bad in (-0.1, 1.5, "x"):
        with pytest.raises(ValueError):
            enc.fidelity_threshold = bad


def test_mps_constructor_errors():
    v = np.ones(4) / 2
    for bad in (-1, 0, 1.5):
        with pytest.raises(ValueError):
            MPS(statevector=v, bond_dimension=bad)
    with pytest.raises(ValueError):
        MPS(statevector=np.array([1.0, 0.0]), bond_dimension=4)       # 1 qubit
    with pytest.raises(TypeError):
        MPS(mps="not an mps", bond_dimension=4)
    with pytest.raises(ValueError):
        MPS(bond_dimension=4)                                           # neither argument
    with pytest.raises(ValueError):
        MPS(statevector=v, mps=DeviceMPS([torch.zeros(1, 2, 1)], None), bond_dimension=4)   # both
    with pytest.raises(ValueError):
        MPS(mps=DeviceMPS([torch.zeros(1, 2, 1, dtype=torch.complex128)], None), bond_dimension=4)  # 1 tensor


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    enc = Sequential(GateListCircuit)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.prepare_state(np.ones(16) / 4, 8, num_layers=1)


def test_num_layers_validation_needs_no_device():
    enc = Sequential(GateListCircuit)

    class Dummy:
        pass
    for bad in (0, -1, 1.5, "2"):
        with pytest.raises(ValueError, match="positive integer"):
            enc.prepare_mps(Dummy(), num_layers=bad)


def test_ket_and_circuit_helpers():
    k = Ket([3, 0, 0, 4j, 0])
    assert k.num_qubits == 3 and abs(np.linalg.norm(k.data) - 1) < 1e-15
    with pytest.raises(ValueError):
        k.change_indexing("diagonal")
    c = GateListCircuit(2)
    c.unitary(np.eye(2), 0)
    c.unitary(np.eye(4), [0, 1])
    assert c.count_ops() == {"unitary1": 1, "unitary2": 1} and c.get_depth() == 2
    with pytest.raises(ValueError):
        c.unitary(np.eye(2), [0, 1])


def test_circuit_emission_convention():
    """sequential.py:181-187: site i -> qubit N-1-i, two-qubit gate on [N-2-i, N-1-i]."""
    n = 3
    gates = np.zeros((1, n, 16), dtype=complex)
    rng = np.random.default_rng(0)
    kinds = [[2, 2, 1]]
    mats = []
    for i, k in enumerate(kinds[0]):
        d = 4 if k == 2 else 2
        q, _ = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
        gates[0, i, : d * d] = q.reshape(-1)
        mats.append(q)
    enc = Sequential(GateListCircuit)
    circ = enc._circuit_from_unitary_layers(n, gates, kinds)
    assert [q for _, q in circ.gates] == [[1, 2], [0, 1], [0]]
    from oracle import qmprs_oracle as O
    layers = [[(0, 2, mats)]]
    assert np.abs(circ.get_statevector() - O.circuit_state(layers, n)).max() < 1e-14


def test_ket_snake_and_compress_hand_computed():
    """base.py:99-102 pre-processing (quick Ket.change_indexing / Ket.compress), hand-computed examples.
    snake: the register read as a 2 x 2^(n-1) image, every second row reversed (boustrophedon);
    compress: the smallest `p` percent of the amplitudes (by modulus) are zeroed, the rest renormalised."""
    v = np.arange(1, 9, dtype=float)                      # 3 qubits: rows [1 2 3 4], [5 6 7 8]
    k = Ket(v)
    k.change_indexing("snake")
    want = np.array([1, 2, 3, 4, 8, 7, 6, 5], dtype=float)
    assert np.allclose(k.data, want / np.linalg.norm(want), atol=1e-15)
    k.change_indexing("snake")                             # an involution
    assert np.allclose(k.data, v / np.linalg.norm(v), atol=1e-15)
    k2 = Ket([1.0, 2.0, 3.0, 4.0])                         # fewer than 3 qubits: unchanged
    k2.change_indexing("snake")
    assert np.allclose(k2.data, np.array([1, 2, 3, 4]) / np.sqrt(30.0))
    k.change_indexing("row")
    assert np.allclose(k.data, v / np.linalg.norm(v), atol=1e-15)
    # compress: 8 amplitudes, 37.5 % -> the 3 smallest moduli (1, -2, 3j) go to zero
    w = np.array([5, 1, -2, 3j, 4, 6, -7, 8j], dtype=complex)
    k3 = Ket(w)
    k3.compress(37.5)
    kept = np.array([5, 0, 0, 0, 4, 6, -7, 8j], dtype=complex)
    assert np.allclose(k3.data, kept / np.linalg.norm(kept), atol=1e-15)
    k4 = Ket(w)
    k4.compress(0.0)
    assert np.allclose(k4.data, w / np.linalg.norm(w), atol=1e-15)


def test_prepare_state_preprocessing_branches_are_applied_in_reference_order():
    """base.py:96-104: Ket wrap -> change_indexing -> compress -> MPS -> prepare_mps, observed through a stub
    encoder (no device needed)."""
    from qmprs_b200.synthesis.mps_encoding.base import MPSEncoder
    import qmprs_b200.synthesis.mps_encoding.base as base_mod
    seen = {}

    class FakeMPS:
        def __init__(self, statevector, bond_dimension):
            seen["data"] = np.array(statevector.data)
            seen["chi"] = bond_dimension

    class Stub(MPSEncoder):
        def prepare_mps(self, mps, **kwargs):
            seen["kwargs"] = kwargs
            return "circuit"

    old = base_mod.MPS
    base_mod.MPS = FakeMPS
    try:
        v = np.array([5, 1, -2, 3, 4, 6, -7, 8], dtype=complex)
        out = Stub(GateListCircuit).prepare_state(v, 16, compression_percentage=25.0, index_type="snake", num_layers=3)
    finally:
        base_mod.MPS = old
    snake = np.array([5, 1, -2, 3, 8, -7, 6, 4], dtype=complex)
    snake[[1, 2]] = 0                                       # 25 % of 8 = the two smallest moduli
    assert out == "circuit" and seen["chi"] == 16 and seen["kwargs"] == {"num_layers": 3}
    assert np.allclose(seen["data"], snake / np.linalg.norm(snake), atol=1e-15)
