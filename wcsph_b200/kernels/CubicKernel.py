"""CubicKernel: import path of the reference (kernels/CubicKernel.py); the class lives in splines.py."""
from .splines import CubicKernel  # noqa: F401
