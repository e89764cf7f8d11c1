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
