"""Drop-in import paths of the reference (``from qmprs.synthesis.mps_encoding import
Sequential``, ``from qmprs.primitives import MPS``) backed by :mod:`qmprs_b200`."""
__all__ = ["primitives", "synthesis"]

from qmprs import primitives, synthesis
