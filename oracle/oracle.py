"""ctypes front end of the CPU oracle (oracle/wcsph_oracle.c).

TEST INFRASTRUCTURE ONLY -- pinned on the executed reference (see oracle/wcsph_oracle.h, tests/test_ref_exec.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  Nothing under wcsph_b200/ does.

The per-solver constant tables below restate the module-level constants of the
reference scripts (file:line cited per entry); they are evaluated in float64
exactly as the Python host code of the reference evaluates them and narrowed to
f32 once, which is what Taichi bakes into its kernels.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "wcsph_oracle.c")
_SRC_BD = os.path.join(_HERE, "boundry_oracle.c")
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp; output oracle/_build/liboracle.so (git-ignored)."""
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    hdr = os.path.join(_HERE, "wcsph_oracle.h")
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= max(os.path.getmtime(_SRC), os.path.getmtime(_SRC_BD), os.path.getmtime(hdr))):
        return _LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall",
           "-o", _LIB, _SRC, _SRC_BD, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


class OracleParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "searchR", "m_k", "m_l", "h3inv", "m_k_raw", "m_l_raw", "coh_m_k", "coh_m_c", "adh_m_k",
        "rho_L0", "rho_S0", "VL0", "VS0", "liqiudMass")] + [("gravity", C.c_float * 3)] + [
        (n, C.c_float) for n in (
        "dim_coff", "viscosity", "viscosity_b", "viscosity_err", "tension_coff", "tension_coff_b",
        "viscosity_omega", "vorticity_coff", "vorticity_init", "stiffness", "pci_coff",
        "omega_relax", "eps", "particleRadius", "user_max_t", "user_min_t")]


def get_pci_coff(particleRadius=0.025, searchR=None, pi=3.1415926):
    """pcisph.py:74-115 (CpuGradW + GetPciCoff), float64 numpy like the reference."""
    gridR = particleRadius * 2.0
    if searchR is None:
        searchR = gridR * 2.0
    h3 = searchR * searchR * searchR
    m_l = 48.0 / (pi * h3)

    def CpuGradW(r):
        res = np.array([0.0, 0.0, 0.0])
        rl = np.linalg.norm(r)
        q = rl / searchR
        if (rl > 1.0e-5) and (q <= 1.0):
            gradq = r / (rl * searchR)
            if q <= 0.5:
                res = m_l * q * (3.0 * q - 2.0) * gradq
            else:
                factor = 1.0 - q
                res = -m_l * (factor * factor) * gradq
        return res

    supportRadius = searchR
    diam = 2.0 * particleRadius
    sumGradW = np.array([0.0, 0.0, 0.0])
    sumGradW2 = 0.0
    V00 = particleRadius * particleRadius * particleRadius * 0.8 * 8.0
    xi = np.array([0.0, 0.0, 0.0])
    xj = np.array([-supportRadius, -supportRadius, -supportRadius])
    while xj[0] <= supportRadius:
        while xj[1] <= supportRadius:
            while xj[2] <= supportRadius:
                r = xi - xj
                dist = np.linalg.norm(r)
                if dist < supportRadius:
                    grad = CpuGradW(r)
                    sumGradW += grad
                    dist_grad = np.linalg.norm(grad)
                    sumGradW2 += dist_grad * dist_grad
                xj[2] += diam
            xj[1] += diam
            xj[2] = -supportRadius
        xj[0] += diam
        xj[1] = -supportRadius
        xj[2] = -supportRadius
    beta = 2.0 * V00 * V00
    dist_sumgrad = np.linalg.norm(sumGradW)
    return 1.0 / (beta * (dist_sumgrad * dist_sumgrad + sumGradW2))


def solver_constants(solver, particleRadius=0.025, **over):
    """Module-level constants of each reference script -> dict of Python floats."""
    R = particleRadius
    VL0 = R * R * R * 0.8 * 8.0                      # ParticleData.py:20, sesph.py:36
    c = dict(particleRadius=R, rho_L0=1000.0, rho_S0=1000.0, VL0=VL0, gravity=(0.0, -9.81, 0.0),
             dim_coff=10.0, viscosity_err=0.05, tension_coff=0.0, tension_coff_b=0.0,
             viscosity_omega=0.1, vorticity_coff=0.01, vorticity_init=0.5, stiffness=50000.0,
             pci_coff=0.0, omega_relax=0.5, eps=1e-5, user_max_t=0.005, user_min_t=0.0001)
    if solver == "dfsph":
        # ParticleData(0.025): HashGrid(particleR*2.0) ParticleData.py:27; kernels use hash_grid.searchR dfsph.py:76
        c.update(hash_gridR=R * 2.0, VS0=VL0, viscosity=10.0, viscosity_b=10.0, style=0, pi=math.pi)
    elif solver == "sesph":
        gridR = R * 2.0                               # sesph.py:25
        c.update(hash_gridR=gridR * 2.0,              # ParticleData(gridR) sesph.py:71 (Q5)
                 VS0=VL0 * 2.0, viscosity=0.1, viscosity_b=0.0, style=1, pi=3.1415926)
    elif solver == "pcisph":
        gridR = R * 2.0
        c.update(hash_gridR=gridR * 2.0, VS0=VL0 * 2.0, viscosity=0.05, viscosity_b=0.0, style=1,
                 pi=3.1415926)
        c["pci_coff"] = get_pci_coff(R)
    elif solver == "iisph":
        gridR = R * 2.0
        c.update(hash_gridR=gridR * 2.0, VS0=VL0, viscosity=2.0, viscosity_b=3.0, style=1,
                 pi=3.1415926, user_min_t=0.00005)
    else:
        raise ValueError(solver)
    c["searchR"] = (R * 2.0) * 2.0                    # physics support: 0.1 in every script
    c["liqiudMass"] = VL0 * c["rho_L0"]
    c.update(over)
    return c


def make_params(c):
    """dict of Python floats -> OracleParams (the single f64->f32 narrowing point)."""
    p = OracleParams()
    h = c["searchR"]
    pi = c["pi"]
    if c["style"] == 0:                               # CubicKernel.py:12-16
        h3 = 1.0 / (h * h * h)
        m_k_raw, m_l_raw = 8.0 / pi, 48.0 / pi
        p.h3inv, p.m_k_raw, p.m_l_raw = h3, m_k_raw, m_l_raw
        p.m_k = m_k_raw * h3
        p.m_l = m_l_raw * h3                          # self.m_l*self.h3 folded in Python
    else:                                             # sesph.py:41-45
        h3 = h * h * h
        p.m_k = 8.0 / (pi * h3)
        p.m_l = 48.0 / (pi * h3)
        p.h3inv, p.m_k_raw, p.m_l_raw = 1.0 / h3, 8.0 / pi, 48.0 / pi
    p.searchR = h
    p.coh_m_k = 32.0 / (math.pi * math.pow(h, 9.0))   # CohesionKernel.py:15
    p.coh_m_c = math.pow(h, 6.0) / 64.0               # CohesionKernel.py:16
    p.adh_m_k = 0.007 / math.pow(h, 3.25)             # AdhesionKernel.py:15
    for k in ("rho_L0", "rho_S0", "VL0", "VS0", "liqiudMass", "dim_coff", "viscosity", "viscosity_b",
              "viscosity_err", "tension_coff", "tension_coff_b", "viscosity_omega", "vorticity_coff",
              "vorticity_init", "stiffness", "pci_coff", "omega_relax", "eps", "particleRadius",
              "user_max_t", "user_min_t"):
        setattr(p, k, c[k])
    p.gravity[0], p.gravity[1], p.gravity[2] = c["gravity"]
    return p


_VEC3 = {"pos", "vel", "vel_guess", "omega", "d_vel", "d_omega", "normal", "cg_r", "cg_dir", "cg_Ad",
         "cg_s", "d_ii", "dij_pj", "pos_star", "vel_star", "d_vel_pre"}
_SCAL = {"vel_max", "pressure", "rho", "adv_rho", "alpha_coff", "kappa", "kappa_v", "a_ii", "pressure_pre"}
_GLOB = {"avg_density_err", "cg_delta", "cg_delta_old", "cg_delta_zero", "rho_err", "deltaT"}

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.POINTER(OracleParams)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_field.restype = C.c_void_p
        L.oracle_field.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_flag.restype = C.c_int
        L.oracle_flag.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_set_params.argtypes = [C.c_void_p, C.POINTER(OracleParams)]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_cubic_W_norm.restype = C.c_float
        L.oracle_cubic_W_norm.argtypes = [C.POINTER(OracleParams), C.c_float, C.c_int]
        L.oracle_cubic_gradW.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_cohesion_W_norm.restype = C.c_float
        L.oracle_cohesion_W_norm.argtypes = [C.POINTER(OracleParams), C.c_float]
        L.oracle_adhesion_W_norm.restype = C.c_float
        L.oracle_adhesion_W_norm.argtypes = [C.POINTER(OracleParams), C.c_float]
        L.oracle_canvas_clear.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_canvas_draw_particle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_mc_update_grid.restype = C.c_int
        L.oracle_mc_update_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_mc_cal_surface_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_double,
                                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_mc_marching_cube.restype = C.c_int
        L.oracle_mc_marching_cube.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_pd_compute_color_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.POINTER(OracleParams), C.c_void_p, C.c_void_p]
        L.oracle_pd_cal_anistropic_kernel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.oracle_mc_cal_surface_point_anistropic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                                             C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def canvas_draw_particle(pos, liquid_count, view, proj, sx, sy, style):
    """clear_canvas + draw_particle (Canvas.py:205-209, dfsph.py:585-593 / sesph.py:201-207) -> (img[sx,sy,3], depth[sx,sy])."""
    pos = np.ascontiguousarray(pos, np.float32)
    view = np.ascontiguousarray(view, np.float32).reshape(16)
    proj = np.ascontiguousarray(proj, np.float32).reshape(16)
    img = np.empty((sx, sy, 3), np.float32)
    depth = np.empty((sx, sy), np.float32)
    L = lib()
    L.oracle_canvas_clear(img.ctypes.data, depth.ctypes.data, sx, sy)
    L.oracle_canvas_draw_particle(pos.ctypes.data, len(pos), int(liquid_count), view.ctypes.data, proj.ctypes.data,
                                  sx, sy, int(style), img.ctypes.data, depth.ctypes.data)
    return img, depth


class Oracle:
    """One reference scene: ParticleData + HashGrid + the solver's module-level fields."""

    def __init__(self, solver, pos, liquid_count, maxInGrid=64, maxNeighbour=2048, threads=1, **over):
        self.solver = solver
        self.c = solver_constants(solver, **over)
        self.params = make_params(self.c)
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.count = int(pos.shape[0])
        self.liquid_count = int(liquid_count)
        self.maxInGrid, self.maxNeighbour = maxInGrid, maxNeighbour
        # ParticleData.py:111-113: bbox over every added point, kept in float32
        self.maxb = pos.max(axis=0).astype(np.float32)
        self.minb = pos.min(axis=0).astype(np.float32)
        L = lib()
        L.oracle_set_threads(threads)
        self.h = L.oracle_create(self.count, self.liquid_count, pos.ctypes.data, self.c["hash_gridR"],
                                 maxInGrid, maxNeighbour, self.maxb.ctypes.data, self.minb.ctypes.data,
                                 C.byref(self.params))
        self.call("reset_param")

    def __del__(self):
        try:
            if self.h:
                lib().oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ptr(self, name):
        p = lib().oracle_field(self.h, name.encode())
        if not p:
            raise KeyError(name)
        return p

    def field(self, name):
        """numpy VIEW of a per-particle field (writes go through)."""
        p = self._ptr(name)
        NL, N = self.liquid_count, self.count
        if name == "pos":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(N, 3))
        if name in _VEC3:
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(NL, 3))
        if name in _SCAL:
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(NL,))
        if name == "cg_Minv":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(NL, 3, 3))
        if name == "gridCount":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(N,))
        if name == "grid":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(N, self.maxInGrid))
        if name == "neighborCount":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(NL,))
        if name == "neighbor":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(NL, self.maxNeighbour))
        if name == "blockSize":
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(3,))
        if name in ("min_boundary", "max_boundary"):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(3,))
        raise KeyError(name)

    def get(self, name):
        assert name in _GLOB
        return float(C.cast(self._ptr(name), C.POINTER(C.c_float))[0])

    def set(self, name, v):
        assert name in _GLOB
        C.cast(self._ptr(name), C.POINTER(C.c_float))[0] = v

    def flag(self, name):
        return lib().oracle_flag(self.h, name.encode())

    def set_constants(self, **over):
        self.c.update(over)
        self.params = make_params(self.c)
        lib().oracle_set_params(self.h, C.byref(self.params))

    def call(self, fn):
        """call `<solver>_<fn>` (or hashgrid_update_grid for fn == 'update_grid')."""
        name = "hashgrid_update_grid" if fn == "update_grid" else "%s_%s" % (self.solver, fn)
        f = getattr(lib(), name)
        f.argtypes = [C.c_void_p]
        f.restype = None
        f(self.h)

    def step(self):
        self.call("step")

    def state(self, names):
        return {n: (self.field(n).copy() if n not in _GLOB else self.get(n)) for n in names}


class McOracle:
    """MCGrid(particleR, maxInGrid, maxNeighbour, particle_data) of ParticleData.py:29 with the boundaries of
    ParticleData.py:177 (scene bbox -+ searchR) -- serial restatement of MarchingCubeGrid.py:160-209,262-352."""

    def __init__(self, pos, liquid_count, particleR=0.025, maxInGrid=4, liqiudMass=None, threads=8):
        self.pos = np.ascontiguousarray(pos, np.float32)
        self.liquid_count = int(liquid_count)
        self.gridR = particleR * 0.9                                   # MarchingCubeGrid.py:22
        self.searchR = self.gridR * 4.0                                # :25
        self.maxInGrid = maxInGrid
        maxb = self.pos.max(axis=0).reshape(1, 3).astype(np.float32) + self.searchR      # ParticleData.py:151-158,177
        minb = self.pos.min(axis=0).reshape(1, 3).astype(np.float32) - self.searchR
        self.maxb, self.minb = maxb.astype(np.float32), minb.astype(np.float32)
        self.block = np.array([int(float(self.maxb[0, k] - self.minb[0, k]) / self.gridR + 1) for k in range(3)], np.int32)   # :61-63
        self.grid_num = int(self.block[0]) * int(self.block[1]) * int(self.block[2])
        if liqiudMass is None:
            c = solver_constants("dfsph", particleR)
            liqiudMass = c["liqiudMass"]
        self.liqiudMass = float(liqiudMass)
        lib().oracle_set_threads(threads)
        self.gridCount = np.zeros(self.grid_num, np.int32)
        self.grid = np.zeros((self.grid_num, maxInGrid), np.int32)
        self.surface_value = np.zeros(self.grid_num, np.float32)

    def update_grid(self, pos=None):
        if pos is not None:
            self.pos = np.ascontiguousarray(pos, np.float32)
        return lib().oracle_mc_update_grid(self.pos.ctypes.data, len(self.pos), self.minb.ctypes.data, self.block.ctypes.data,
                                           self.gridR, self.maxInGrid, self.gridCount.ctypes.data, self.grid.ctypes.data)

    def cal_surface_point(self, rho):
        rho = np.ascontiguousarray(rho, np.float32)
        lib().oracle_mc_cal_surface_point(self.pos.ctypes.data, rho.ctypes.data, self.liquid_count, self.liqiudMass,
                                          self.minb.ctypes.data, self.block.ctypes.data, self.gridR, self.maxInGrid,
                                          self.gridCount.ctypes.data, self.grid.ctypes.data, self.surface_value.ctypes.data)
        return self.surface_value

    def marching_cube(self, edgetable, tritable, surface_value=None, max_vertex=3000000):
        sv = self.surface_value if surface_value is None else np.ascontiguousarray(surface_value, np.float32)
        e = np.ascontiguousarray(edgetable, np.int32)
        t = np.ascontiguousarray(tritable, np.int32)
        tri = np.zeros((max_vertex, 3), np.float32)
        n = lib().oracle_mc_marching_cube(sv.ctypes.data, self.minb.ctypes.data, self.block.ctypes.data, self.gridR,
                                          e.ctypes.data, t.ctypes.data, tri.ctypes.data, max_vertex)
        return n, tri[:min(n, max_vertex)]

    def cal_surface_point_anistropic(self, rho, pos_avr, G):
        """MarchingCubeGrid.py:215-243 (restatement only: the GPU engine does not build the anisotropic branch yet)."""
        rho = np.ascontiguousarray(rho, np.float32)
        pa = np.ascontiguousarray(pos_avr, np.float32)
        G = np.ascontiguousarray(G, np.float32)
        lib().oracle_mc_cal_surface_point_anistropic(self.pos.ctypes.data, pa.ctypes.data, G.ctypes.data, rho.ctypes.data, self.liquid_count,
                                                     self.liqiudMass, self.minb.ctypes.data, self.block.ctypes.data, self.gridR, self.maxInGrid,
                                                     self.gridCount.ctypes.data, self.grid.ctypes.data, self.surface_value.ctypes.data)
        return self.surface_value


def compute_color_map(o):
    """ParticleData.compute_color_map (ParticleData.py:188-218) on an Oracle whose update_grid / density are current
    -> (color[NL], color_grad[NL,3]).  Restatement only."""
    NL = o.liquid_count
    color, grad = np.zeros(NL, np.float32), np.zeros((NL, 3), np.float32)
    pos, rho = np.ascontiguousarray(o.field("pos")), np.ascontiguousarray(o.field("rho"))
    lib().oracle_pd_compute_color_map(pos.ctypes.data, rho.ctypes.data, NL, o.field("neighborCount").ctypes.data,
                                      o.field("neighbor").ctypes.data, o.maxNeighbour, C.byref(o.params), color.ctypes.data, grad.ctypes.data)
    return color, grad


def cal_anistropic_kernel(o, mc_searchR):
    """ParticleData.cal_anistropic_kernel (ParticleData.py:223-285) -> (pos_avr[NL,3], G[NL,3,3]).  Restatement only."""
    NL = o.liquid_count
    pos_avr, G = np.zeros((NL, 3), np.float32), np.zeros((NL, 3, 3), np.float32)
    pos = np.ascontiguousarray(o.field("pos"))
    lib().oracle_pd_cal_anistropic_kernel(pos.ctypes.data, NL, o.field("neighborCount").ctypes.data, o.field("neighbor").ctypes.data,
                                          o.maxNeighbour, float(mc_searchR), pos_avr.ctypes.data, G.ctypes.data)
    return pos_avr, G


# ---- boundry.py (Poisson-disk boundary sampler), restatement in oracle/boundry_oracle.c ---------------------------
class BdParams(C.Structure):
    _fields_ = [("n", C.c_int), ("padding", C.c_int), ("hash_size", C.c_int), ("phase_vec_max", C.c_int),
                ("radius", C.c_float), ("gridR", C.c_float), ("minp", C.c_float * 3)]


class BoundryOracle:
    """boundry.py:160-457 from an injected initial point set (init_pos [n,3] f32, init_id [n] i32, face ids)."""
    SAMPLE_CAP = 5                                           # hash_sample_size, boundry.py:61

    def __init__(self, tri_normal, init_pos, init_id, min_point, particleRadius=0.025):
        import math
        n = len(init_pos)
        m = 1
        while m < n:
            m <<= 1
        self.p = BdParams()
        self.p.n, self.p.padding = n, (m >> 1) << 1          # get_pot_num(n) << 1, boundry.py:82-86,166
        self.p.hash_size, self.p.phase_vec_max = n * 3, n // 8          # :167-168
        self.p.radius, self.p.gridR = particleRadius, particleRadius / math.sqrt(3.0)      # :21-22
        for k in range(3):
            self.p.minp[k] = float(min_point[k])
        P = self.p.padding
        self.tri_normal = np.ascontiguousarray(tri_normal, np.float32)
        self.pos = np.zeros((P, 3), np.float32); self.pos[:n] = init_pos
        self.id = np.zeros(P, np.int32); self.id[:n] = init_id
        self.cell = np.zeros((P, 3), np.int32)
        self.start_index = np.zeros(self.p.hash_size, np.int32)
        self.hcell = np.zeros((self.p.hash_size, 3), np.int32)
        self.hash_trace = np.zeros(n, np.int32)
        self.phase_group_count = np.zeros(27, np.int32)
        self.phase_group = np.zeros((27, max(self.p.phase_vec_max, 1), 3), np.int32)
        self.sample_count = np.zeros(self.p.hash_size, np.int32)
        self.sample = np.zeros((self.p.hash_size, self.SAMPLE_CAP), np.int32)
        self.possion_sample = np.zeros((n, 3), np.float32)
        self.selected = np.zeros(n, np.int32)
        self.n_sample = 0
        self.L = lib()

    def cells(self):
        self.L.oracle_bd_cells(C.byref(self.p), self.pos.ctypes.data_as(C.c_void_p), self.cell.ctypes.data_as(C.c_void_p))

    def sort(self):
        self.L.oracle_bd_bitonic_sort(C.byref(self.p), self.cell.ctypes.data_as(C.c_void_p), self.pos.ctypes.data_as(C.c_void_p),
                                      self.id.ctypes.data_as(C.c_void_p))

    def build_hmap(self):
        self.L.oracle_bd_build_hmap.restype = C.c_int
        return self.L.oracle_bd_build_hmap(C.byref(self.p), self.cell.ctypes.data_as(C.c_void_p), self.start_index.ctypes.data_as(C.c_void_p),
                                           self.hcell.ctypes.data_as(C.c_void_p), self.hash_trace.ctypes.data_as(C.c_void_p),
                                           self.phase_group_count.ctypes.data_as(C.c_void_p), self.phase_group.ctypes.data_as(C.c_void_p))

    def sample_launch(self, pg, trial):
        f = self.L.oracle_bd_sample_launch
        f.restype = C.c_int
        V = C.c_void_p
        self.n_sample = f(C.byref(self.p), C.c_int(pg), C.c_int(trial), C.c_int(min(int(self.phase_group_count[pg]), self.p.phase_vec_max)),
                          self.phase_group.ctypes.data_as(V), self.cell.ctypes.data_as(V), self.pos.ctypes.data_as(V), self.id.ctypes.data_as(V),
                          self.tri_normal.ctypes.data_as(V), self.start_index.ctypes.data_as(V), self.sample_count.ctypes.data_as(V),
                          self.sample.ctypes.data_as(V), C.c_int(self.SAMPLE_CAP), self.possion_sample.ctypes.data_as(V),
                          self.selected.ctypes.data_as(V), C.c_int(self.n_sample))
        return self.n_sample

    @staticmethod
    def launch_order(trial_total=10, phases=27):
        """the (phase, trial) sequence of the reference's main loop, boundry.py:421-457: phase_process is incremented BEFORE the
        first launch, so trial 0 never visits phase group 0"""
        out, phase, trial = [], 0, 0
        while True:
            if trial < trial_total:
                phase += 1
                if phase % phases == 0:
                    trial += 1
                    phase = 0
            if trial < trial_total:
                out.append((phase, trial))
            else:
                return out

    def run(self):
        self.cells(); self.sort(); self.build_hmap()
        for pg, trial in self.launch_order():
            self.sample_launch(pg, trial)
        return self.possion_sample[:self.n_sample]
