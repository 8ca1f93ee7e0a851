"""Runs a few steps of every solver on its as-shipped 8k scene (stepwise + fused); meant to be launched
under `compute-sanitizer --tool memcheck` / `--tool racecheck` (the reference has none)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
from wcsph_b200 import scenes
for solver in ("sesph", "pcisph", "iisph", "dfsph"):
    pts, nl = getattr(scenes, "scene_" + solver)()
    m = importlib.import_module("wcsph_b200." + solver)
    m.init_scene(pts, nl)
    m.reset_param()
    for _ in range(2):
        m.step()
    m.step_fused(2)
    m.particle_data.pos.to_numpy()
    # the rows next to the path: canvas pass and surface reconstruction
    cv = m.sph_canvas
    cv.static_cam(0.0, 1.0, 0.0)
    cv.clear_canvas()
    m.draw_particle()
    lit = int((cv.img.to_numpy()[:, :, 0] > 0).sum())
    g = m.particle_data.mc_grid
    g.update_grid()
    g.cal_surface_point()
    nv = g.marching_cube()
    print(solver, "ok, status", m.particle_data.hash_grid.status(), "canvas pixels", lit, "mesh vertices", nv)
