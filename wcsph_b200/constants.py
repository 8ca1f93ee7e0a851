"""Kernel constants the reference bakes in at JIT time, evaluated like its Python host code.

Each solver script of the reference keeps its own module-level constants (sesph.py:24-62,
pcisph.py:24-69, iisph.py:25-92) except dfsph.py, which reads them off ParticleData
(ParticleData.py:13-22,61-87).  `solver_params()` turns such a namespace into the POD
`wcsph_params` the CUDA library takes: float64 arithmetic in Python, narrowed to f32 once
at the ctypes boundary.
"""
import math

from . import _lib


def cubic_constants(searchR, style, pi):
    """CubicKernel.py:12-16 (style 0) or sesph.py:41-45 (style 1) -> (m_k, m_l, m_k_raw, h3inv)."""
    if style == 0:
        h3 = 1.0 / (searchR * searchR * searchR)
        m_k, m_l = 8.0 / pi, 48.0 / pi
        return m_k * h3, m_l * h3, m_k, h3
    h3 = searchR * searchR * searchR
    return 8.0 / (pi * h3), 48.0 / (pi * h3), 8.0 / pi, 1.0 / h3


def solver_params(ns):
    """ns: mapping with the reference's constant names -> _lib.Params."""
    g = ns.get
    p = _lib.Params()
    h = g("searchR")
    style = int(g("kernel_style", 0))
    pi = g("pi", math.pi)
    p.searchR = h
    p.m_k, p.m_l, p.m_k_raw, p.h3inv = cubic_constants(h, style, pi)
    p.kernel_style = style
    p.coh_m_k = 32.0 / (math.pi * math.pow(h, 9.0))       # CohesionKernel.py:15
    p.coh_m_c = math.pow(h, 6.0) / 64.0                   # CohesionKernel.py:16
    p.adh_m_k = 0.007 / math.pow(h, 3.25)                 # AdhesionKernel.py:15
    p.rho_L0 = g("rho_L0", 1000.0)
    p.rho_S0 = g("rho_S0", p.rho_L0)
    p.VL0 = g("VL0")
    p.VS0 = g("VS0", p.VL0)
    p.liqiudMass = g("liqiudMass", g("VL0") * g("rho_L0", 1000.0))
    gr = g("gravity", (0.0, -9.81, 0.0))
    p.gravity[0], p.gravity[1], p.gravity[2] = float(gr[0]), float(gr[1]), float(gr[2])
    p.dim_coff = g("dim_coff", 10.0)
    p.viscosity = g("viscosity", 0.0)
    p.viscosity_b = g("viscosity_b", 0.0)
    p.viscosity_err = g("viscosity_err", 0.05)
    p.tension_coff = g("tension_coff", 0.0)
    p.tension_coff_b = g("tension_coff_b", 0.0)
    p.viscosity_omega = g("viscosity_omega", 0.1)
    p.vorticity_coff = g("vorticity_coff", 0.01)
    p.vorticity_init = g("vorticity_init", 0.5)
    p.stiffness = g("stiffness", 50000.0)
    p.pci_coff = g("pci_coff", 0.0)
    p.omega_relax = g("omega", g("omega_relax", 0.5))
    p.eps = g("eps", 1e-5)
    p.particleRadius = g("particleRadius", 0.025)
    p.user_max_t = g("user_max_t", 0.005)
    p.user_min_t = g("user_min_t", 0.0001)
    return p
