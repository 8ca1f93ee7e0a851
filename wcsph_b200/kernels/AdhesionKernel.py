"""AdhesionKernel: import path of the reference (kernels/AdhesionKernel.py); the class lives in splines.py."""
from .splines import AdhesionKernel  # noqa: F401
