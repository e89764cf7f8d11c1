__all__ = ["MPS"]

from qmprs_b200.primitives.mps import MPS
