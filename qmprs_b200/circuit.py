"""Duck-typed circuit backend.

The reference emits gates through ``circuit_framework(num_qubits)`` / ``.unitary(matrix,
qubit_indices)`` (qmprs/synthesis/mps_encoding/sequential.py:182-187, 208) and nothing
else, so any class with those members works (quick's QiskitCircuit in the reference).
``GateListCircuit`` records the gates; ``get_statevector`` is a small host-side
little-endian simulator for verification, outside the accelerated path.
"""
from __future__ import annotations

import numpy as np


class GateListCircuit:
    def __init__(self, num_qubits: int) -> None:
        self.num_qubits = int(num_qubits)
        self.gates: list[tuple[np.ndarray, list[int]]] = []

    def unitary(self, matrix, qubit_indices) -> None:
        q = [int(qubit_indices)] if np.isscalar(qubit_indices) else [int(x) for x in qubit_indices]
        m = np.array(matrix, dtype=np.complex128)
        if m.shape != (2 ** len(q), 2 ** len(q)):
            raise ValueError("matrix shape does not match the number of qubits")
        self.gates.append((m, q))

    def count_ops(self) -> dict:
        out = {"unitary1": 0, "unitary2": 0}
        for _, q in self.gates:
            out["unitary%d" % len(q)] += 1
        return out

    def get_depth(self) -> int:
        """Depth counted in emitted one-/two-qubit unitaries (not quick's U3/CX basis)."""
        level = [0] * self.num_qubits
        for _, q in self.gates:
            d = max(level[x] for x in q) + 1
            for x in q:
                level[x] = d
        return max(level) if level else 0

    def get_statevector(self) -> np.ndarray:
        n = self.num_qubits
        psi = np.zeros(2 ** n, dtype=np.complex128)
        psi[0] = 1.0
        for m, q in self.gates:
            t = psi.reshape([2] * n)
            if len(q) == 2:
                qa, qb = q                      # matrix index = 2*bit(qb) + bit(qa)
                axb, axa = n - 1 - qb, n - 1 - qa
                t = np.tensordot(m.reshape(2, 2, 2, 2), t, axes=([2, 3], [axb, axa]))
                t = np.moveaxis(t, [0, 1], [axb, axa])
            else:
                ax = n - 1 - q[0]
                t = np.moveaxis(np.tensordot(m, t, axes=([1], [ax])), 0, ax)
            psi = np.ascontiguousarray(t).reshape(-1)
        return psi
