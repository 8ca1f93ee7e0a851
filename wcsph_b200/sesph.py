"""sesph -- drop-in for the reference's sesph.py (state-equation SPH, Tait gamma = 7).

Module constants as sesph.py:24-62; `init_particle`, the five former @ti.kernels as
zero-argument functions (sesph.py:131-196), `step()` = one pass of sesph.py:220-225.
Nothing runs at import; the GUI loop is out of scope.
"""

from .ParticleData import ParticleData
from .Canvas import Canvas
from . import scenes

current_time = 0.0
eps = 1e-5
imgSize = 512           # sesph.py:15

# particle param (sesph.py:24-38)
particleRadius = 0.025
gridR = particleRadius * 2.0
searchR = gridR * 2.0
invGridR = 1.0 / gridR
particleDimX = 20
particleDimY = 20
particleDimZ = 20
particleLiquidNum = particleDimX * particleDimY * particleDimZ
boundary = 2.0
rho_0 = 1000.0
VL0 = particleRadius * particleRadius * particleRadius * 0.8 * 8.0
VS0 = VL0 * 2.0
liqiudMass = VL0 * rho_0

# kernel param (sesph.py:41-45)
pi = 3.1415926
h3 = searchR * searchR * searchR
m_k = 8.0 / (pi * h3)
m_l = 48.0 / (pi * h3)

gravity = (0.0, -9.81, 0.0)
stiffness = 50000.0
dim_coff = 10.0
viscosity = 0.1
viscosity_b = 0.0

particle_data = None
vel = d_vel = rho = pressure = deltaT = None


def _namespace():
    return dict(searchR=searchR, kernel_style=1, pi=pi, rho_L0=rho_0, rho_S0=rho_0, VL0=VL0, VS0=VS0,
                liqiudMass=liqiudMass, gravity=gravity, dim_coff=dim_coff, viscosity=viscosity,
                viscosity_b=viscosity_b, stiffness=stiffness, eps=eps, particleRadius=particleRadius)


def _bind(pd):
    global particle_data, vel, d_vel, rho, pressure, deltaT, particleLiquidNum
    particle_data = pd
    sph_canvas.bind(pd)
    particleLiquidNum = pd.liquid_count
    pd.setup_data_gpu()
    pd.setup_data_cpu()
    vel, d_vel, rho, pressure, deltaT = pd.vel, pd.d_vel, pd.rho, pd.pressure, pd.deltaT


def init_particle(filename=None, **kw):
    """sesph.py:66-92 (the filename argument is unused there too)."""
    pts, nl = scenes.scene_sesph(particleRadius, (particleDimX, particleDimY, particleDimZ))
    init_scene(pts, nl, **kw)


def init_scene(points, liquid_count, **kw):
    pd = ParticleData(gridR, solver="sesph", constants=None, **kw)     # ParticleData(gridR): sesph.py:71 (Q5)
    pd._namespace = _namespace
    pd.add_liquid_points(points[:liquid_count])
    pd.add_solid_points(points[liquid_count:])
    _bind(pd)


def _k(name):
    particle_data.call("sesph_" + name)


def reset_param(): _k("reset_param")
def update_advection_density(): _k("update_advection_density")
def update_pressure(): _k("update_pressure")
def compute_force(): _k("compute_force")
def integrator_sesph(): _k("integrator_sesph")


def step():
    """sesph.py:220-225 + :234-235."""
    global current_time
    particle_data.hash_grid.update_grid()
    update_advection_density()
    update_pressure()
    compute_force()
    integrator_sesph()
    dt = deltaT.to_numpy()[0]
    current_time += dt
    particle_data.check()          # raise if the device dropped pairs (the reference only prints, HashGrid.py:73,103)
    return dt


def step_fused(n=1):
    particle_data.call("sesph_step", int(n))


def main(steps=100):
    init_particle("boundry.obj")
    reset_param()
    for _ in range(steps):
        step()
        print("time:%.3f" % current_time, "step:%.4f" % deltaT.to_numpy()[0])


sph_canvas = Canvas(imgSize, imgSize)        # sesph.py:213 (host object only; device buffers appear on first use)


def draw_particle():
    """sesph.py:201-207: liquids as 3-pixel circle outlines, solids as grey points -- one launch."""
    sph_canvas.draw_particle(particle_data, style=0)


if __name__ == "__main__":
    main()
