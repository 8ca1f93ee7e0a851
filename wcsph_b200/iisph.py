"""iisph -- drop-in for the reference's iisph.py (implicit incompressible SPH + Weiler-2018
implicit viscosity).

Module constants as iisph.py:25-92, `init_particle`, host helpers `compute_nonpressure_force`
(iisph.py:114-126) and `solve_pressure` (iisph.py:130-139), the former @ti.kernels as
zero-argument functions (iisph.py:178-396), `step()` = iisph.py:419-427.
"""

from .ParticleData import ParticleData
from .Canvas import Canvas
from . import scenes

current_time = 0.0
eps = 1e-5
imgSizeX = 512          # iisph.py:15-16
imgSizeY = 512
test_id = 0

particleRadius = 0.025
gridR = particleRadius * 2.0
invGridR = 1.0 / gridR
particleDimX = 20
particleDimY = 20
particleDimZ = 20
particleLiquidNum = particleDimX * particleDimY * particleDimZ

rho_L0 = 1000.0
rho_S0 = rho_L0
VL0 = particleRadius * particleRadius * particleRadius * 0.8 * 8.0
VS0 = VL0
liqiudMass = VL0 * rho_L0
boundary = 2.0

searchR = gridR * 2.0
pi = 3.1415926
h3 = searchR * searchR * searchR
m_k = 8.0 / (pi * h3)
m_l = 48.0 / (pi * h3)

gravity = (0.0, -9.81, 0.0)
vs_iter = 0
dv_iter = 0
pr_iter = 0
user_max_t = 0.005
user_min_t = 0.00005

dim_coff = 10.0
omega = 0.5
viscosity = 2.0
viscosity_b = 3.0
viscosity_err = 0.05

particle_data = None
vel_guess = vel = vel_max = d_vel = a_ii = d_ii = dij_pj = pressure_pre = pressure = rho = adv_rho = None
avg_density_err = cg_delta = cg_delta_old = cg_delta_zero = deltaT = None
cg_Minv = cg_r = cg_dir = cg_Ad = cg_s = None


def _namespace():
    return dict(searchR=searchR, kernel_style=1, pi=pi, rho_L0=rho_L0, rho_S0=rho_S0, VL0=VL0, VS0=VS0,
                liqiudMass=liqiudMass, gravity=gravity, dim_coff=dim_coff, viscosity=viscosity,
                viscosity_b=viscosity_b, viscosity_err=viscosity_err, omega=omega, eps=eps,
                particleRadius=particleRadius, user_max_t=user_max_t, user_min_t=user_min_t)


def _bind(pd):
    g = globals()
    g["particle_data"] = pd
    sph_canvas.bind(pd)
    g["particleLiquidNum"] = pd.liquid_count
    pd.setup_data_gpu()
    pd.setup_data_cpu()
    for n in ("vel_guess", "vel", "vel_max", "d_vel", "a_ii", "d_ii", "dij_pj", "pressure_pre", "pressure", "rho", "adv_rho",
              "avg_density_err", "cg_delta", "cg_delta_old", "cg_delta_zero", "deltaT", "cg_Minv", "cg_r", "cg_dir", "cg_Ad", "cg_s"):
        g[n] = getattr(pd, n)


def init_particle(filename="box_boundry", **kw):
    """iisph.py:99-112."""
    pts, nl = scenes.scene_iisph(filename, particleRadius, (particleDimX, particleDimY, particleDimZ))
    init_scene(pts, nl, **kw)


def init_scene(points, liquid_count, **kw):
    pd = ParticleData(gridR, solver="iisph", **kw)                     # iisph.py:101 (Q5)
    pd._namespace = _namespace
    pd.add_liquid_points(points[:liquid_count])
    pd.add_solid_points(points[liquid_count:])
    _bind(pd)


def _k(name):
    particle_data.call("iisph_" + name)


def reset_param(): _k("reset_param")
def init_viscosity_para(): _k("init_viscosity_para")
def compute_viscosity_force(): _k("compute_viscosity_force")
def compute_density(): _k("compute_density")
def combine_nonpressure(): _k("combine_nonpressure")
def compute_advection(): _k("compute_advection")
def update_iter_info(): _k("update_iter_info")
def update_pressure_force(): _k("update_pressure_force")
def update_pos(): _k("update_pos")


def compute_nonpressure_force():
    """iisph.py:114-126."""
    global vs_iter
    init_viscosity_para()
    vs_iter = 0
    while vs_iter < 100:
        compute_viscosity_force()
        vs_iter += 1
        if cg_delta[0] <= viscosity_err * cg_delta_zero[0] or cg_delta_zero[0] < eps:
            break
    combine_nonpressure()


def solve_pressure():
    """iisph.py:130-139."""
    global pr_iter
    pr_iter = 0
    err = 0.0
    while (err > 0.001 or pr_iter < 2) and (pr_iter < 100):
        update_iter_info()
        update_pressure_force()
        err = avg_density_err.to_numpy()[0] / float(particleLiquidNum)
        pr_iter += 1


def step():
    """iisph.py:419-427 + :434-435."""
    global current_time
    particle_data.hash_grid.update_grid()
    compute_density()
    compute_nonpressure_force()
    compute_advection()
    solve_pressure()
    update_pos()
    dt = deltaT.to_numpy()[0]
    current_time += dt
    particle_data.check()          # raise if the device dropped pairs (the reference only prints, HashGrid.py:73,103)
    return dt


def step_fused(n=1):
    global vs_iter, pr_iter
    particle_data.call("iisph_step", int(n))
    vs_iter, _, pr_iter = particle_data.iters()


def main(steps=100, filename="box_boundry"):
    init_particle(filename)
    reset_param()
    for _ in range(steps):
        step()
        print("time:%.3f" % current_time, "step:%.4f" % deltaT.to_numpy()[0], "viscorcity:", vs_iter, "pressure:", pr_iter)


sph_canvas = Canvas(imgSizeX, imgSizeY)        # iisph.py:410 (host object only; device buffers appear on first use)


def draw_particle():
    """iisph.py:401-406: liquids as 3-pixel circle outlines, solids as grey points -- one launch."""
    sph_canvas.draw_particle(particle_data, style=0)


if __name__ == "__main__":
    main()
