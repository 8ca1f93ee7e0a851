import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from wcsph_b200 import pcisph, scenes, _lib
dims = (60, 60, 60)
pts, nl = scenes.dam_break(*dims)
pcisph.init_scene(pts, nl)
pcisph.reset_param()
pcisph.set_tension(0.1, 0.05)
for s in range(14):
    pcisph.step_fused(1)
    pd = pcisph.particle_data
    nc = pd.hash_grid.neighborCount.to_numpy()
    pos = pd.pos.to_numpy()[:nl]; vel = pd.vel.to_numpy(); rho = pd.rho.to_numpy(); dv = pd.d_vel.to_numpy(); pr = pd.pressure.to_numpy()
    print(s, "flags", pd.hash_grid.status(), "nc max", nc.max(), "pr_iter", pcisph.pr_iter, "|v|max %.3g |dvel|max %.3g rho [%.1f, %.1f] p max %.3g nan %s" % (
        np.abs(vel).max(), np.abs(dv).max(), rho.min(), rho.max(), np.abs(pr).max(), np.isnan(pos).any()))
