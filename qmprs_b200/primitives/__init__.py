__all__ = ["MPS", "DeviceMPS", "GateTensor"]

from qmprs_b200.primitives.mps import MPS, DeviceMPS, GateTensor
