__all__ = ["MPSEncoder", "Sequential"]

from qmprs_b200.synthesis.mps_encoding.base import MPSEncoder
from qmprs_b200.synthesis.mps_encoding.sequential import Sequential
