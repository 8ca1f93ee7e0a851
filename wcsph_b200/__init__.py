"""wcsph_b200 -- B200-native engine for the per-step SPH hot path of lyd405121/wcsph.

Host side mirrors the reference's module surface (ParticleData, HashGrid, kernels.*, Canvas, MarchingCubeGrid,
sesph / pcisph / iisph / dfsph); the compute is hand-written sm_100a CUDA behind the
C ABI of include/wcsph_b200.h (libwcsph_b200.so, loaded with ctypes).  There is no CPU
fallback: the solver modules raise if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
