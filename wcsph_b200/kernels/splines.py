"""Host-side mirrors of the reference's three smoothing-kernel classes (kernels/CubicKernel.py:12-54,
CohesionKernel.py:12-33, AdhesionKernel.py:12-33).

What the solvers consume from these classes is their CONSTANTS (`searchR`, `h3`, `m_k`, `m_l`, `m_c`):
the kernel bodies are inlined into the CUDA sweeps (csrc/engine.cuh `cubic_W` / `cubic_gradW`,
csrc/tension.cuh `coh_W` / `adh_W`).  The evaluators below exist for inspection and for the closed-form
identity tests; they are float32 numpy, accept scalars or arrays, and are not on the hot path.
"""
import math

import numpy as np

_f32 = np.float32


def _as_r(v):
    """|v| of a 3-vector (or a stack of them) in float32."""
    a = np.asarray(v, dtype=np.float32)
    return np.sqrt(np.sum(a * a, axis=-1, dtype=np.float32))


class _Kernel:
    """shared plumbing: `Cubic_W(vec)` is always `Cubic_W_norm(|vec|)` in the reference"""

    def __init__(self, searchR):
        self.searchR = searchR

    def Cubic_W(self, vec):
        return self.Cubic_W_norm(_as_r(vec))


class CubicKernel(_Kernel):
    """W(q) = (8/pi h^3) {6q^3 - 6q^2 + 1 | 2(1-q)^3 | 0}, gradW = (48/pi h^3) {q(3q-2) | -(1-q)^2} r/(|r| h)."""

    def __init__(self, searchR):
        super().__init__(searchR)
        self.h3 = 1.0 / (searchR * searchR * searchR)
        self.m_k = 8.0 / math.pi
        self.m_l = 48.0 / math.pi

    def Cubic_W_P(self, q):
        q = np.asarray(q, dtype=np.float32)
        inner = _f32(6.0) * q * q * q - _f32(6.0) * q * q + _f32(1.0)
        outer = _f32(2.0) * (_f32(1.0) - q) ** 3
        return np.where(q <= 0.5, inner, np.where(q <= 1.0, outer, _f32(0.0))).astype(np.float32)

    def Cubic_W_norm(self, r):
        q = np.asarray(r, dtype=np.float32) / _f32(self.searchR)
        return self.Cubic_W_P(q) * _f32(self.m_k) * _f32(self.h3)

    def CubicGradW(self, vec):
        v = np.asarray(vec, dtype=np.float32)
        rl = _as_r(v)
        q = rl / _f32(self.searchR)
        c = _f32(self.m_l * self.h3)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(q <= 0.5, c * q * (_f32(3.0) * q - _f32(2.0)), -c * (_f32(1.0) - q) ** 2) / (rl * _f32(self.searchR))
        s = np.where((rl > 1.0e-5) & (q <= 1.0), s, _f32(0.0)).astype(np.float32)
        return (v * s[..., None]).astype(np.float32) if v.ndim > 1 else (v * s).astype(np.float32)


class CohesionKernel(_Kernel):
    """Akinci 2013 cohesion spline as the reference writes it: m_k (h-r)^3 r^3 for r > h/2, else
    2 m_k (h-r)^3 r^3 - h^6/64 (the constant is NOT scaled by m_k, CohesionKernel.py:27)."""

    def __init__(self, searchR):
        super().__init__(searchR)
        self.m_k = 32.0 / (math.pi * math.pow(searchR, 9.0))
        self.m_c = math.pow(searchR, 6.0) / 64.0

    def Cubic_W_norm(self, r):
        r = np.asarray(r, dtype=np.float32)
        h = _f32(self.searchR)
        core = _f32(self.m_k) * (h - r) ** 3 * r ** 3
        val = np.where(r > _f32(0.5) * h, core, _f32(2.0) * core - _f32(self.m_c))
        return np.where(r * r <= h * h, val, _f32(0.0)).astype(np.float32)


class AdhesionKernel(_Kernel):
    """Akinci 2013 adhesion spline: (0.007 / h^3.25) (-4 r^2/h + 6 r - 2 h)^(1/4) on h/2 < r <= h."""

    def __init__(self, searchR):
        super().__init__(searchR)
        self.m_k = 0.007 / math.pow(searchR, 3.25)

    def Cubic_W_norm(self, r):
        r = np.asarray(r, dtype=np.float32)
        h = _f32(self.searchR)
        rad = np.maximum(_f32(-4.0) * r * r / h + _f32(6.0) * r - _f32(2.0) * h, _f32(0.0))
        val = _f32(self.m_k) * np.power(rad, _f32(0.25))
        return np.where((r * r <= h * h) & (r > _f32(0.5) * h), val, _f32(0.0)).astype(np.float32)
