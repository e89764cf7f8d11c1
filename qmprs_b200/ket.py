"""Minimal stand-in for ``quick.primitives.Ket`` (quick-core is not installable offline).

Only what the hot path touches (qmprs/synthesis/mps_encoding/base.py:96-102,
qmprs/primitives/mps.py:170-179, 270): wrap an array-like as a normalised, power-of-two
padded column of amplitudes exposing ``data``, ``num_qubits``, ``change_indexing`` and
``compress``.  This is host-side input preparation, outside the accelerated path.
"""
from __future__ import annotations

import numpy as np


class Ket:
    def __init__(self, data):
        if isinstance(data, Ket):
            data = data.data
        v = np.array(data, dtype=np.complex128).reshape(-1)
        if v.size == 0:
            raise ValueError("The statevector must not be empty.")
        n = max(1, int(np.ceil(np.log2(v.size))))
        if v.size != 2 ** n:
            v = np.concatenate([v, np.zeros(2 ** n - v.size, dtype=np.complex128)])
        nrm = np.linalg.norm(v)
        if nrm == 0:
            raise ValueError("The statevector must not be the zero vector.")
        self.norm_scale = float(nrm)
        self.data = v / nrm
        self.num_qubits = n

    def change_indexing(self, index_type: str) -> None:
        """"row" keeps the order; "snake" reverses every other row of the 2 x (2^n/2)
        reshape (quick's image convention; recalled, not on any BASELINE config)."""
        if index_type == "row":
            return
        if index_type == "snake":
            if self.num_qubits >= 3:
                m = self.data.reshape(2, -1).copy()
                m[1::2, :] = m[1::2, ::-1]
                self.data = m.reshape(-1)
            return
        raise ValueError("Index type not supported.")

    def compress(self, compression_percentage: float) -> None:
        """Zero the smallest ``compression_percentage`` % amplitudes and renormalise."""
        flat = self.data.copy()
        k = int(flat.size * compression_percentage / 100.0)
        if k > 0:
            idx = np.argsort(np.abs(flat), kind="stable")[:k]
            flat[idx] = 0
            nrm = np.linalg.norm(flat)
            if nrm > 0:
                flat = flat / nrm
        self.data = flat

    def __len__(self):
        return self.data.size
