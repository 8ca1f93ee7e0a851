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
    if solver == "dfsph":
        # the anisotropic branch of the reconstruction (N2, second half): colour map, covariance + eigen-decomposition, anisotropic field
        pd = m.particle_data
        pd.compute_color_map()
        pd.cal_anistropic_kernel()
        g.update_grid()
        g.cal_surface_point_anistropic()
        print("anisotropic branch ok, mesh vertices", g.marching_cube(), "status", pd.hash_grid.status())

# boundry.py (N3): Poisson-disk sampling of a small closed box mesh, all stages
import tempfile
from wcsph_b200 import boundry
work = tempfile.mkdtemp(prefix="sanitize_boundry_")
v = [(-0.11, 0.0, -0.08), (0.13, 0.0, -0.08), (0.13, 0.17, -0.08), (-0.11, 0.17, -0.08),
     (-0.11, 0.0, 0.12), (0.13, 0.0, 0.12), (0.13, 0.17, 0.12), (-0.11, 0.17, 0.12)]
f = [(1, 3, 2), (1, 4, 3), (5, 6, 7), (5, 7, 8), (1, 2, 6), (1, 6, 5), (4, 7, 3), (4, 8, 7), (1, 5, 8), (1, 8, 4), (2, 3, 7), (2, 7, 6)]
with open(os.path.join(work, "box.obj"), "w") as fo:
    for p in v:
        fo.write("v %.6f %.6f %.6f\n" % p)
    for t in f:
        fo.write("f %d %d %d\n" % t)
samples = boundry.main(os.path.join(work, "box"), seed=3)
print("boundry ok,", boundry.numInitialPoints, "initial points ->", len(samples), "samples")
