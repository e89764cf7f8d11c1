__all__ = ["mps_encoding"]

from qmprs_b200.synthesis import mps_encoding
