"""pcisph -- drop-in for the reference's pcisph.py (predictive-corrective incompressible SPH).

Module constants as pcisph.py:24-69, host-side `CpuGradW` / `GetPciCoff` (pcisph.py:74-115,
float64 numpy like the reference), `init_particle`, the former @ti.kernels as zero-argument
functions (pcisph.py:194-285), `sovel_pressure` (pcisph.py:147-157, reference spelling) and
`step()` = pcisph.py:307-311.  D-PCI: compute_nonpressure_force is two-phase.
"""
import numpy as np

from .ParticleData import ParticleData
from .Canvas import Canvas
from . import scenes

current_time = 0.0
eps = 1e-5
imgSize = 512           # pcisph.py:15
test_id = 0

particleRadius = 0.025
gridR = particleRadius * 2.0
searchR = gridR * 2.0
invGridR = 1.0 / gridR
boundary = 2.0
particleDimX = 20
particleDimY = 20
particleDimZ = 20
particleLiquidNum = particleDimX * particleDimY * particleDimZ

rho_0 = 1000.0
VL0 = particleRadius * particleRadius * particleRadius * 0.8 * 8.0
VS0 = VL0 * 2.0
liqiudMass = VL0 * rho_0

pi = 3.1415926
h3 = searchR * searchR * searchR
m_k = 8.0 / (pi * h3)
m_l = 48.0 / (pi * h3)

gravity = (0.0, -9.81, 0.0)
dim_coff = 10.0
viscosity = 0.05
viscosity_b = 0.0
tension_coff = 0.0        # config 3 (Akinci tension on PCISPH): see DESIGN.md, not in the reference script
tension_coff_b = 0.0
pr_iter = 0
pci_coff = None

particle_data = None
vel = d_vel = d_vel_pre = pos_star = vel_star = rho = adv_rho = pressure = rho_err = deltaT = None


def CpuGradW(r):
    """pcisph.py:74-85: float64 cubic-spline gradient; `r` may be one vector or an (n, 3) stack."""
    r = np.asarray(r, dtype=np.float64)
    rl = np.sqrt((r * r).sum(axis=-1))
    q = rl / searchR
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(q <= 0.5, m_l * q * (3.0 * q - 2.0), -m_l * (1.0 - q) * (1.0 - q)) / (rl * searchR)
    f = np.where((rl > 1.0e-5) & (q <= 1.0), f, 0.0)
    return r * f[..., None] if r.ndim > 1 else r * f


def GetPciCoff():
    """pcisph.py:87-115: the PCISPH delta for a filled lattice of spacing 2R around the origin,
    1 / (2 V0^2 (|sum gradW|^2 + sum |gradW|^2)).  The reference walks the lattice by repeated `+= diam`;
    cumsum reproduces those running sums exactly, so the lattice nodes are the same doubles."""
    diam = 2.0 * particleRadius
    n = int(2.0 * searchR / diam) + 3
    axis = np.cumsum(np.concatenate([[-searchR], np.full(n, diam)]))
    axis = axis[axis <= searchR]
    X, Y, Z = np.meshgrid(axis, axis, axis, indexing="ij")
    r = -np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)          # xi - xj with xi = 0
    inside = np.sqrt((r * r).sum(axis=1)) < searchR
    g = CpuGradW(r[inside])
    V00 = particleRadius * particleRadius * particleRadius * 0.8 * 8.0
    s = g.sum(axis=0)
    return 1.0 / (2.0 * V00 * V00 * (float(s @ s) + float((g * g).sum())))


def _namespace():
    return dict(searchR=searchR, kernel_style=1, pi=pi, rho_L0=rho_0, rho_S0=rho_0, VL0=VL0, VS0=VS0,
                liqiudMass=liqiudMass, gravity=gravity, dim_coff=dim_coff, viscosity=viscosity,
                viscosity_b=viscosity_b, eps=eps, particleRadius=particleRadius,
                pci_coff=pci_coff if pci_coff is not None else GetPciCoff(),
                tension_coff=tension_coff, tension_coff_b=tension_coff_b)


def _bind(pd):
    global particle_data, vel, d_vel, d_vel_pre, pos_star, vel_star, rho, adv_rho, pressure, rho_err, deltaT
    global particleLiquidNum, pci_coff
    if pci_coff is None:
        pci_coff = GetPciCoff()                     # pcisph.py:299
    particle_data = pd
    sph_canvas.bind(pd)
    particleLiquidNum = pd.liquid_count
    pd.setup_data_gpu()
    pd.setup_data_cpu()
    vel, d_vel, d_vel_pre, pos_star, vel_star = pd.vel, pd.d_vel, pd.d_vel_pre, pd.pos_star, pd.vel_star
    rho, adv_rho, pressure, rho_err, deltaT = pd.rho, pd.adv_rho, pd.pressure, pd.rho_err, pd.deltaT


def init_particle(filename=None, **kw):
    """pcisph.py:117-143."""
    pts, nl = scenes.scene_pcisph(particleRadius, (particleDimX, particleDimY, particleDimZ))
    init_scene(pts, nl, **kw)


def init_scene(points, liquid_count, **kw):
    # predict_density evaluates gradW(pos_i - pos_star_j) (pcisph.py:266-268): list candidates a
    # little beyond h so a neighbour that the prediction moves into range is not missed
    kw.setdefault("cull_scale", 1.25)
    kw.setdefault("list_cap_liquid", 128)      # (1.25)^3 x the ~32..40 in-range neighbours, + alias duplicates
    kw.setdefault("list_cap_solid", 128)
    pd = ParticleData(gridR, solver="pcisph", **kw)                    # pcisph.py:122 (Q5)
    pd._namespace = _namespace
    pd.add_liquid_points(points[:liquid_count])
    pd.add_solid_points(points[liquid_count:])
    _bind(pd)


def _k(name):
    particle_data.call("pcisph_" + name)


def reset_param(): _k("reset_param")
def compute_nonpressure_force(): _k("compute_nonpressure_force")


def compute_tension():
    """BASELINE configs[2] extension: dfsph.py:265-304 (D-TENSION) added to the PCISPH non-pressure
    acceleration; a no-op on d_vel while tension_coff == tension_coff_b == 0 (as shipped)."""
    _k("compute_tension")


def set_tension(coff, coff_b=0.0):
    global tension_coff, tension_coff_b
    tension_coff, tension_coff_b = float(coff), float(coff_b)
    particle_data.update_params()
def init_iter_info(): _k("init_iter_info")
def update_iter_info(): _k("update_iter_info")
def predict_density(): _k("predict_density")
def update_pos(): _k("update_pos")


def sovel_pressure():
    """pcisph.py:147-157."""
    global pr_iter
    pr_iter = 0
    err = 0.0
    init_iter_info()
    while (err > 0.01 or pr_iter < 3) and (pr_iter < 50):
        update_iter_info()
        predict_density()
        err = rho_err.to_numpy()[0] / float(particleLiquidNum)
        pr_iter += 1


def step():
    """pcisph.py:307-311 + :321-322."""
    global current_time
    particle_data.hash_grid.update_grid()
    compute_nonpressure_force()
    if tension_coff != 0.0 or tension_coff_b != 0.0:
        compute_tension()
    sovel_pressure()
    update_pos()
    dt = deltaT.to_numpy()[0]
    current_time += dt
    particle_data.check()          # raise if the device dropped pairs (the reference only prints, HashGrid.py:73,103)
    return dt


def step_fused(n=1):
    global pr_iter
    particle_data.call("pcisph_step", int(n))
    pr_iter = particle_data.iters()[2]


def main(steps=100):
    init_particle("boundry.obj")
    reset_param()
    for _ in range(steps):
        step()
        print("time:%.3f" % current_time, "step:%.4f" % deltaT.to_numpy()[0], "pressure:", pr_iter)


sph_canvas = Canvas(imgSize, imgSize)        # pcisph.py:296 (host object only; device buffers appear on first use)


def draw_particle():
    """pcisph.py:288-293: liquids as 3-pixel circle outlines, solids as grey points -- one launch."""
    sph_canvas.draw_particle(particle_data, style=0)


if __name__ == "__main__":
    main()
