"""CohesionKernel: import path of the reference (kernels/CohesionKernel.py); the class lives in splines.py."""
from .splines import CohesionKernel  # noqa: F401
