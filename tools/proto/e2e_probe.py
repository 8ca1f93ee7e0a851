import sys, time
sys.path.insert(0, '.')
import torch
from wcsph_b200 import dfsph, scenes, _lib
pts, nl = scenes.dam_break(100, 100, 100)
dfsph.init_scene(pts, nl); dfsph.reset_param()
pd = dfsph.particle_data; L = _lib.load(); ctx = pd._ctx
N = len(pts)
dfsph.step_fused(5); pd.sync()
pos_h = torch.empty((N, 3), dtype=torch.float32).pin_memory(); vel_h = torch.empty((nl, 3), dtype=torch.float32).pin_memory()
pos_h.copy_(torch.from_numpy(pd.pos.to_numpy())); vel_h.copy_(torch.from_numpy(pd.vel.to_numpy()))
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(name, fn, reps=10):
    pd.sync(); fn(); pd.sync()
    e0, e1 = ev(), ev(); t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); pd.sync(); t1 = time.perf_counter()
    print("%-28s gpu %.3f ms  wall %.3f ms" % (name, e0.elapsed_time(e1) / reps, (t1 - t0) * 1e3 / reps))
timeit("set pos (12 MB H2D+scatter)", lambda: _lib.check(L.wcsph_field_set_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4)))
timeit("set vel", lambda: _lib.check(L.wcsph_field_set_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4)))
timeit("get pos (13.6 MB)", lambda: _lib.check(L.wcsph_field_get_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4)))
timeit("get vel", lambda: _lib.check(L.wcsph_field_get_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4)))
timeit("step", lambda: dfsph.step_fused(1, fetch_iters=False))
d = torch.empty(12_000_000, dtype=torch.uint8, device="cuda"); h = torch.empty(12_000_000, dtype=torch.uint8).pin_memory()
timeit("raw H2D 12 MB", lambda: d.copy_(h, non_blocking=True))
timeit("raw D2H 12 MB", lambda: h.copy_(d, non_blocking=True))
def full():
    _lib.check(L.wcsph_field_set_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
    _lib.check(L.wcsph_field_set_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
    dfsph.step_fused(1, fetch_iters=False)
    _lib.check(L.wcsph_field_get_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
    _lib.check(L.wcsph_field_get_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
    pd.sync()
timeit("full e2e step", full)
