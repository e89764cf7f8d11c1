"""``MPSEncoder`` base class (reference: qmprs/synthesis/mps_encoding/base.py:31-131)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Literal

from qmprs_b200.ket import Ket
from qmprs_b200.primitives.mps import MPS

__all__ = ["MPSEncoder"]


class MPSEncoder(ABC):
    """Approximate state preparation through a matrix product state.

    ``circuit_framework`` is any class with ``circuit_framework(num_qubits)``,
    ``.num_qubits`` and ``.unitary(matrix, qubit_indices)`` (quick's ``Circuit``
    subclasses in the reference).
    """

    def __init__(self, circuit_framework) -> None:
        self.circuit_framework = circuit_framework

    def prepare_state(self, statevector, bond_dimension: int, compression_percentage: float = 0.0,
                      index_type: Literal["row", "snake"] = "row", **kwargs):
        """base.py:58-106: wrap in a Ket, re-index, optionally compress, build the MPS
        and dispatch to :meth:`prepare_mps`."""
        if not isinstance(statevector, Ket):
            statevector = Ket(statevector)
        statevector.change_indexing(index_type)
        if compression_percentage > 0.0:
            statevector.compress(compression_percentage)
        mps = MPS(statevector, bond_dimension=bond_dimension)
        return self.prepare_mps(mps, **kwargs)

    @abstractmethod
    def prepare_mps(self, mps: MPS, **kwargs):
        """Prepare the quantum state from an :class:`MPS`."""
