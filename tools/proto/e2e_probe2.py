import sys, time
sys.path.insert(0, '.')
import torch
from wcsph_b200 import dfsph as mod, scenes, _lib
pts, nl = scenes.dam_break(100, 100, 100)
mod.init_scene(pts, nl); mod.reset_param()
pd = mod.particle_data; L = _lib.load(); ctx = pd._ctx
N = len(pts)
fused = lambda n: mod.step_fused(n, fetch_iters=False)
stage = sys.argv[1] if len(sys.argv) > 1 else "all"
fused(5); pd.sync()
pd.launch_count(reset=True)
fused(20); pd.sync()
if stage in ("all", "iters"): iters = mod.iters_log(20)
if stage in ("all", "lc"): launches = pd.launch_count()
if stage in ("all", "check"): pd.check(); flags = pd.hash_grid.status()
pos_h = torch.empty((N, 3), dtype=torch.float32).pin_memory(); vel_h = torch.empty((nl, 3), dtype=torch.float32).pin_memory()
pos_h.copy_(torch.from_numpy(pd.pos.to_numpy())); vel_h.copy_(torch.from_numpy(pd.vel.to_numpy()))
def e2e_step():
    _lib.check(L.wcsph_field_set_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
    _lib.check(L.wcsph_field_set_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
    fused(1)
    _lib.check(L.wcsph_field_get_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
    _lib.check(L.wcsph_field_get_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
    pd.sync()
for _ in range(2): e2e_step()
torch.cuda.synchronize()
per = []
for _ in range(20):
    t0 = time.perf_counter(); e2e_step(); per.append((time.perf_counter() - t0) * 1e3)
print(stage, "per-step wall ms:", " ".join("%.2f" % x for x in per))
