__all__ = ["mps_encoding"]

from qmprs.synthesis import mps_encoding
