"""qmprs_b200: B200-native (sm_100a) implementation of the qmprs MPS hot path.

Public surface mirrors the reference: ``qmprs_b200.primitives.MPS``,
``qmprs_b200.synthesis.mps_encoding.{MPSEncoder, Sequential}`` (also importable under
the reference's own paths through the ``qmprs`` shim package).  Importing the package
does not touch CUDA; the first computation loads ``libqmprs_b200.so`` and raises if it
or the GPU is missing (no CPU fallback).
"""
__all__ = ["primitives", "synthesis", "Ket", "GateListCircuit", "U3CXCircuit"]

from qmprs_b200.ket import Ket
from qmprs_b200.circuit import GateListCircuit
from qmprs_b200.transpile import U3CXCircuit
from qmprs_b200 import primitives, synthesis
