"""dfsph -- drop-in for the reference's dfsph.py (module surface of SURVEY.md 8b).

Same module constants, globals (`particle_data`, `vs_iter`, `dv_iter`, `pr_iter`, `deltaT`,
`alpha_coff`, `kappa`, `kappa_v`), zero-argument kernel functions and host helpers
(`init_particle`, `compute_nonpressure_force`, `optimize_time_step`, `solve_vel_divergence`,
`solve_pressure`) with the reference's control flow (dfsph.py:59-164); each former
@ti.kernel is one call into libwcsph_b200.  `step()` is one pass of the reference main loop
(dfsph.py:606-617) through those functions; `step_fused(n)` runs the same sequence inside
the library with the convergence loops evaluated from device scalars.

Nothing here runs at import: call `init_particle(...)`, `reset_param()`, then `step()`.
`sph_canvas` / `draw_particle()` are the reference's canvas pass; only the ti.GUI window is left out.
"""
import math

import numpy as np

from . import _lib
from .ParticleData import ParticleData
from .Canvas import Canvas
from .kernels.CubicKernel import CubicKernel
from .kernels.CohesionKernel import CohesionKernel
from .kernels.AdhesionKernel import AdhesionKernel

# gui param (dfsph.py:19-24)
current_time = 0.0
total_time = 5.0
eps = 1e-5
imgSizeX = 512          # dfsph.py:19-20
imgSizeY = 512
test_id = 0

# particle param (dfsph.py:28-32)
particleRadius = 0.025
particleDimX = 20
particleDimY = 20
particleDimZ = 20
particleLiquidNum = particleDimX * particleDimY * particleDimZ

# CFL time step (dfsph.py:36-42)
vs_iter = 0
dv_iter = 0
pr_iter = 0
user_max_t = 0.005
user_min_t = 0.0001

deltaT = None        # 1-element field, bound by init_particle
alpha_coff = None    # dfsph.py:46-48, bound by init_particle
kappa = None
kappa_v = None

particle_data = None
kernel_c = None
kernel_adh = None
kernel_coh = None


def _bind(pd):
    global particle_data, deltaT, alpha_coff, kappa, kappa_v, kernel_c, kernel_adh, kernel_coh, particleLiquidNum
    particle_data = pd
    sph_canvas.bind(pd)
    particleLiquidNum = pd.liquid_count
    kernel_c = CubicKernel(pd.hash_grid.searchR)          # dfsph.py:76-78
    kernel_adh = AdhesionKernel(pd.hash_grid.searchR)
    kernel_coh = CohesionKernel(pd.hash_grid.searchR)
    pd.setup_data_gpu()
    pd.setup_data_cpu()
    deltaT = pd.deltaT
    alpha_coff, kappa, kappa_v = pd.alpha_coff, pd.kappa, pd.kappa_v


def _make_pd(**kw):
    pd = ParticleData(particleRadius, solver="dfsph", **kw)
    ns = pd._namespace
    def namespace():
        d = ns()
        d.update(eps=eps, particleRadius=particleRadius, user_max_t=user_max_t, user_min_t=user_min_t)
        return d
    pd._namespace = namespace
    return pd


def init_particle(filename, **kw):
    """dfsph.py:59-82."""
    pd = _make_pd(**kw)
    ZxY = particleDimZ * particleDimY
    dis = particleRadius * 2.0
    n = particleDimX * particleDimY * particleDimZ
    i = np.arange(n)
    pts = np.stack([(i // ZxY - particleDimX / 2).astype(np.float64) * dis + dis * 0.5,
                    ((i % ZxY) // particleDimZ).astype(np.float64) * dis + 0.2,
                    (i % particleDimZ - particleDimZ / 2).astype(np.float64) * dis + dis * 0.5], axis=1)
    pd.add_liquid_points(pts)
    pd.add_obj(filename)
    _bind(pd)


def init_scene(points, liquid_count, **kw):
    """any scene (e.g. scenes.dam_break) through the same ParticleData calls."""
    pd = _make_pd(**kw)
    pd.add_liquid_points(points[:liquid_count])
    pd.add_solid_points(points[liquid_count:])
    _bind(pd)


def _k(name):
    particle_data.call("dfsph_" + name)


# former @ti.kernels (dfsph.py:168-580)
def reset_param(): _k("reset_param")
def init_viscosity_para(): _k("init_viscosity_para")
def compute_viscosity_force(): _k("compute_viscosity_force")
def compute_density(): _k("compute_density")
def compute_tension(): _k("compute_tension")
def compute_vorticity(): _k("compute_vorticity")
def clear_nonpressure(): _k("clear_nonpressure")
def end_viscosity(): _k("end_viscosity")
def compute_dfsph_coff(): _k("compute_dfsph_coff")
def warmstart_divergence_vel(): _k("warmstart_divergence_vel")
def begin_divergence_iter(): _k("begin_divergence_iter")
def divergence_iter(): _k("divergence_iter")
def end_divergence_iter(): _k("end_divergence_iter")
def warmstart_pressure(): _k("warmstart_pressure")
def begin_pressure_iter(): _k("begin_pressure_iter")
def pressure_iter(): _k("pressure_iter")
def end_pressure_iter(): _k("end_pressure_iter")
def update_vel(): _k("update_vel")
def update_pos(): _k("update_pos")


def cfl_time_step(index):
    """dfsph.py:556-568.  The reference launches this ceil(log2 NL) times as a stride-doubling
    max tree (racy / out of bounds for non-power-of-two NL, Q15); here the first call
    (index == 1) runs one max reduction that leaves the true maximum in vel_max[0] and the
    later calls of the tree are no-ops."""
    if index == 1:
        _k("cfl_max")


def compute_nonpressure_force():
    """dfsph.py:84-103."""
    global vs_iter
    clear_nonpressure()
    compute_tension()
    init_viscosity_para()
    vs_iter = 0
    while vs_iter < 100:
        compute_viscosity_force()
        vs_iter += 1
        if particle_data.cg_delta[0] <= particle_data.viscosity_err * particle_data.cg_delta_zero[0] or particle_data.cg_delta_zero[0] < eps:
            break
    end_viscosity()
    compute_vorticity()


def optimize_time_step():
    """dfsph.py:107-129."""
    size = 1
    while size < particleLiquidNum:
        cfl_time_step(size)
        size = size * 2
    deltaT_np = deltaT.to_numpy()
    vel_max0 = float(particle_data.vel_max0.to_numpy()[0])
    if vel_max0 > eps:
        cfl_factor = 0.5
        time_step = cfl_factor * 0.4 * particleRadius * 2.0 / math.sqrt(vel_max0)
        time_step = min(time_step, user_max_t)
        time_step = max(time_step, user_min_t)
        iter = max(vs_iter, max(pr_iter, vs_iter))
        d = float(deltaT_np[0])
        if iter > 10:
            d = float(np.float32(d * 0.9))
        elif iter < 5:
            d = float(np.float32(d * 1.1))
        d = min(d, time_step)
        deltaT_np[0] = d
        deltaT.from_numpy(deltaT_np)


def solve_vel_divergence():
    """dfsph.py:131-146."""
    global dv_iter
    dv_iter = 0
    warmstart_divergence_vel()
    err = -0.1
    begin_divergence_iter()
    deltaT_np = deltaT.to_numpy()
    while (particle_data.avg_density_err.to_numpy()[0] > err) and (dv_iter < 10):
        divergence_iter()
        err = 0.001 * float(particleLiquidNum) / deltaT_np[0]
        dv_iter += 1
    end_divergence_iter()


def solve_pressure():
    """dfsph.py:150-164."""
    global pr_iter
    warmstart_pressure()
    pr_iter = 0
    err = 0.0
    begin_pressure_iter()
    while (err > 0.001 or pr_iter < 2) and (pr_iter < 100):
        pressure_iter()
        err = particle_data.avg_density_err.to_numpy()[0] / float(particleLiquidNum)
        pr_iter += 1
    end_pressure_iter()


def step():
    """one pass of the reference main loop body, dfsph.py:606-617 + :626-627."""
    global current_time
    particle_data.hash_grid.update_grid()
    compute_density()
    compute_dfsph_coff()
    solve_vel_divergence()
    compute_nonpressure_force()
    optimize_time_step()
    update_vel()
    solve_pressure()
    update_pos()
    dt = deltaT.to_numpy()[0]
    current_time += dt
    # the fused path reads the counters of dfsph.py:122 from the context: keep them current (a later step_fused
    # must see THIS step's pr_iter, Q17)
    _lib.check(_lib.load().wcsph_set_iters(particle_data._ctx, int(vs_iter), int(dv_iter), int(pr_iter)))
    particle_data.check()          # raise if the device dropped pairs (the reference only prints, HashGrid.py:73,103)
    return dt


def step_fused(n=1, fetch_iters=True):
    """n steps inside libwcsph_b200 (wcsph_dfsph_step: one CUDA graph launch per step, loops as
    device-evaluated WHILE nodes).  fetch_iters=True synchronises and updates vs_iter / dv_iter /
    pr_iter like the reference's per-step console line; False leaves the steps queued."""
    global vs_iter, dv_iter, pr_iter
    particle_data.call("dfsph_step", int(n))
    if fetch_iters:
        vs_iter, dv_iter, pr_iter = particle_data.iters()


def iters_log(max_steps=4096):
    """(vs, dv, pr) of the most recent fused steps, oldest first."""
    import ctypes as C
    from . import _lib
    out = (C.c_int * (3 * max_steps))()
    n = C.c_int()
    _lib.check(_lib.load().wcsph_iters_log(particle_data._ctx, out, max_steps, C.byref(n)))
    return [(out[3 * k], out[3 * k + 1], out[3 * k + 2]) for k in range(n.value)]


def set_graph(on=True):
    import ctypes as C
    from . import _lib
    _lib.check(_lib.load().wcsph_set_option(particle_data._ctx, b"graph", 1 if on else 0))


def log_line():
    """dfsph.py:629."""
    dt = deltaT.to_numpy()[0]
    return "time:%.3f step:%.4f viscorcity: %d divergence: %d particle_data.pressure: %d" % (current_time, dt, vs_iter, dv_iter, pr_iter)


def main(steps=100, filename="box_boundry", png_every=0, surface=False):
    """the reference's `while gui.running` loop (dfsph.py:595-646) without the window: step, draw into the canvas
    (dfsph.py:604,621-622), print the console line (:629); `png_every` = n writes every n-th frame as <frame>.png
    (Canvas.export_png_frame), `surface` runs mc_grid.export_surface(current_time) (the line commented out at :635)."""
    init_particle(filename)
    reset_param()
    for k in range(steps):
        sph_canvas.static_cam(0.0, 1.0, 0.0)
        step()
        sph_canvas.clear_canvas()
        draw_particle()
        print(log_line())
        if png_every and k % png_every == 0:
            sph_canvas.export_png_frame(k + 1)
        if surface:
            particle_data.mc_grid.export_surface(current_time)
        if math.isnan(particle_data.pos.to_numpy()[test_id, 0]) or current_time >= total_time:
            break


sph_canvas = Canvas(imgSizeX, imgSizeY)        # dfsph.py:596 (host object only; device buffers appear on first use)


def draw_particle():
    """dfsph.py:585-593: liquids as 3-pixel circle outlines, then a grey point for every particle -- one launch."""
    sph_canvas.draw_particle(particle_data, style=1)


if __name__ == "__main__":
    main()
