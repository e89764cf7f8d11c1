"""``Sequential`` encoder (reference: qmprs/synthesis/mps_encoding/sequential.py:33-600).

Analytic disentangling layers (Ran 2020) followed by environment-tensor optimisation
sweeps (Rudolph et al. 2022), all numerics on the GPU through :mod:`qmprs_b200.host`.
"""
from __future__ import annotations

import numpy as np

from qmprs_b200 import host
from qmprs_b200.primitives.mps import MPS
from qmprs_b200.synthesis.mps_encoding.base import MPSEncoder

__all__ = ["Sequential"]


class Sequential(MPSEncoder):
    GRAPH_CACHE_MAX = 4

    def __init__(self, circuit_framework) -> None:
        super().__init__(circuit_framework)
        self._fidelity_threshold = 1 - 1e-6          # sequential.py:120
        self.last_result = None                      # diagnostics of the last call (not in the reference)
        # "svd": re-split every two-site gate application by a truncated SVD as the reference does
        # (mps.py:968-971).  "exact": gauge-free trivial re-split, identical circuit, no SVD
        # (qmprs_b200.host.apply_inverse_layer); opt-in.
        self.gate_split = "svd"
        # "DallOall": what the reference runs (sequential.py:543-586).  "IterDiOall" / "IterDiOi": the two schedules
        # it names as future work (notebook :459, docstring :410, 428-432) -- qmprs_b200.host.disentangle_iterative.
        # Also accepted per call as the keyword ``schedule=`` of prepare_state / prepare_mps.
        self.schedule = "DallOall"
        # CUDA-graph replay for small registers (qmprs_b200.graphs), OPT-IN: False (default) always runs the
        # eager path; "auto" captures the pipeline the second time the same (n, chi, layers, sweeps) is
        # requested with n <= 16; True forces it.  The replayed pipeline uses fixed Jacobi sweep budgets and
        # no split-K (different summation order), so its gates agree with the eager ones to ~1e-9, not bit
        # for bit -- hence not the default.  At most GRAPH_CACHE_MAX configurations keep their graph.
        self.use_cuda_graphs = False
        self._graph_cache = {}
        self._graph_seen = {}

    @property
    def fidelity_threshold(self) -> float:
        return self._fidelity_threshold

    @fidelity_threshold.setter
    def fidelity_threshold(self, threshold: float) -> None:
        if not isinstance(threshold, (int, float)) or threshold < 0 or threshold > 1:
            raise ValueError("The fidelity threshold must be a float between 0 and 1.")
        self._fidelity_threshold = threshold

    @staticmethod
    def _apply_unitary_layer_to_circuit(circuit, gates_layer, kinds) -> None:
        """sequential.py:155-187: MPS site i <-> circuit qubit N-1-i."""
        n = circuit.num_qubits
        for index, kind in enumerate(kinds):
            if kind == 1:
                circuit.unitary(gates_layer[index, :4].reshape(2, 2).copy(), abs(index - n + 1))
            else:
                circuit.unitary(gates_layer[index].reshape(4, 4).copy(),
                                [abs(index - n + 2), abs(index - n + 1)])

    def _circuit_from_unitary_layers(self, num_sites, gates, kinds_per_layer):
        """sequential.py:189-213."""
        circuit = self.circuit_framework(num_sites)
        for layer_gates, kinds in zip(gates, kinds_per_layer):
            Sequential._apply_unitary_layer_to_circuit(circuit, layer_gates, kinds)
        return circuit

    def _sequential_unitary_circuit(self, mps: MPS, num_layers: int, num_sweeps: int = 0, schedule=None):
        """sequential.py:543-586."""
        schedule = self.schedule if schedule is None else schedule
        if schedule not in host.SCHEDULES:
            raise ValueError("`schedule` must be one of %s." % (host.SCHEDULES,))
        K = mps.mps.K
        A = mps.mps.tensors
        N = mps.num_sites
        record = {}
        if self.gate_split not in ("svd", "exact"):
            raise ValueError("`gate_split` must be 'svd' or 'exact'.")
        # sequential.py:360-376 is skipped only for a right-canonical MPS whose bonds already went through
        # the reference's cutoff (what MPS.from_statevector / MPS.compress return); any other gauge
        # (left-canonical, after apply_unitary_layer, built from raw arrays) is pre-conditioned in full
        pre = mps.mps.form == "right" and getattr(mps.mps, "trimmed", False)
        if schedule != "DallOall":
            gates_all, layer_kinds, overlaps = host.disentangle_iterative(
                K, A, num_layers, num_sweeps, self._fidelity_threshold, schedule, record, split=self.gate_split,
                preconditioned=pre)
            num_sweeps = 0                                 # done inside, per layer
        else:
            gates_all, layer_kinds, overlaps = host.disentangle(K, A, num_layers, self._fidelity_threshold, record,
                                                                split=self.gate_split, preconditioned=pre)
        if num_sweeps > 0:
            target = host.to_dense(K, A)
            host.optimize_layers(K, target, gates_all, layer_kinds, N, num_sweeps)
        L = len(layer_kinds)
        gates = K.to_host(gates_all).reshape(L, N, 16)
        self.last_result = {"gates": gates, "kinds": layer_kinds, "n_layers": L, "overlaps": overlaps,
                            "gates_device": gates_all}
        return self._circuit_from_unitary_layers(N, gates, layer_kinds)

    def prepare_state(self, statevector, bond_dimension, compression_percentage=0.0, index_type="row", **kwargs):
        """base.py:58-106.  Same contract; small registers requested repeatedly are served by a captured
        CUDA graph (static shapes validated on the device, eager re-run when an assumption fails)."""
        num_layers = kwargs.get("num_layers", 1)
        num_sweeps = kwargs.get("num_sweeps", 0)
        if not isinstance(num_layers, int) or num_layers < 1:
            raise ValueError("The number of layers must be a positive integer.")
        from qmprs_b200.ket import Ket
        if not isinstance(statevector, Ket):
            statevector = Ket(statevector)
        n = statevector.num_qubits
        key = (n, int(bond_dimension), num_layers, int(num_sweeps), float(self._fidelity_threshold), self.gate_split)
        want = self.use_cuda_graphs
        plain = (compression_percentage == 0.0 and index_type == "row" and n >= 2
                 and kwargs.get("schedule", self.schedule) == "DallOall")
        if want and plain and (want is True or (n <= 16 and self._graph_seen.get(key, 0) >= 1)):
            prep = self._graph_cache.get(key)
            if prep is None:
                try:
                    from qmprs_b200.graphs import GraphedPreparer
                    prep = GraphedPreparer(n, int(bond_dimension), num_layers, int(num_sweeps),
                                           float(self._fidelity_threshold), lanes=1, split=self.gate_split)
                except Exception:                          # capture not possible here: stay on the eager path
                    prep = False
                while len(self._graph_cache) >= self.GRAPH_CACHE_MAX:      # oldest first (dicts keep insertion order)
                    self._graph_cache.pop(next(iter(self._graph_cache)))
                self._graph_cache[key] = prep
            if prep:
                res = prep.run(np.asarray(statevector.data, dtype=np.complex128).reshape(1, -1))[0]
                self.last_result = {"gates": res["gates"], "kinds": res["kinds"], "n_layers": res["n_layers"],
                                    "overlaps": res.get("overlaps"), "graph": True}
                return self._circuit_from_unitary_layers(n, res["gates"], res["kinds"])
        self._graph_seen[key] = self._graph_seen.get(key, 0) + 1
        return super().prepare_state(statevector, bond_dimension, compression_percentage, index_type, **kwargs)

    def prepare_mps(self, mps: MPS, **kwargs):
        num_layers = kwargs.get("num_layers", 1)
        num_sweeps = kwargs.get("num_sweeps", 0)
        if not isinstance(num_layers, int) or num_layers < 1:
            raise ValueError("The number of layers must be a positive integer.")
        return self._sequential_unitary_circuit(mps, num_layers, num_sweeps, kwargs.get("schedule"))
