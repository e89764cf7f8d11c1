"""qmprs_b200: B200-native (sm_100a) implementation of the qmprs MPS hot path.

Public surface mirrors the reference: ``qmprs_b200.primitives.MPS``,
``qmprs_b200.synthesis.mps_encoding.{MPSEncoder, Sequential}`` (also importable under
the reference's own paths through the ``qmprs`` shim package).  Importing the package
does not touch CUDA; the first computation loads ``libqmprs_b200.so`` and raises if it
or the GPU is missing (no CPU fallback).
"""
import os as _os

# Batches of small states run as concurrent CUDA-graph lanes, one stream each (graphs.py).  The driver
# maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8, maximum 32) when the
# context is created; lanes that share a queue serialise (measured: 8 lanes 109, 16 lanes 206, 32 lanes
# 380 twelve-qubit states/s with 32 queues; no gain beyond 8 lanes with the default).  Respect a user setting.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__all__ = ["primitives", "synthesis", "Ket", "GateListCircuit", "U3CXCircuit"]

from qmprs_b200.ket import Ket
from qmprs_b200.circuit import GateListCircuit
from qmprs_b200.transpile import U3CXCircuit
from qmprs_b200 import primitives, synthesis
