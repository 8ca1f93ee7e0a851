#!/usr/bin/env python
"""Timing of the surface reconstruction on the 1M dam-break scene (35 M grid nodes) via the library's
CUDA-event profiler (the CPU restatement is timed by tests/test_mc_gpu.py -s: only tests/ and bench.py may execute
oracle/).  Usage: python tools/mc_bench.py [reps]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from wcsph_b200 import _lib, dfsph, scenes  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pts, nl = scenes.dam_break(100, 100, 100, jitter=True, config_id=2)
dfsph.init_scene(pts, nl)
dfsph.reset_param()
dfsph.step_fused(3)
g = dfsph.particle_data.mc_grid


def frame():
    g.update_grid()
    g.cal_surface_point()
    return g.marching_cube()


n = frame()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    frame()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
L = _lib.load()
ctx = dfsph.particle_data._ctx
_lib.check(L.wcsph_profile(ctx, 1))
for _ in range(reps):
    frame()
buf = C.create_string_buffer(1 << 16)
_lib.check(L.wcsph_profile_report(ctx, buf, len(buf)))
_lib.check(L.wcsph_profile(ctx, 0))
print("MC grid %s = %d nodes, %d liquids: %d vertices (%d triangles), %.2f ms per reconstruction (update_grid + cal_surface_point + marching_cube)" % (
    tuple(int(x) for x in g.blocknp[0]), g.grid_num, nl, n, n // 3, ms))
for line in buf.value.decode().splitlines():
    name, c, t = line.split("\t")
    per = float(t) / int(c)
    extra = ""
    if name == "k_mc_surface":
        extra = "  %.1f G nodes/s; %.0f GB/s of the 4 B/node result stream" % (g.grid_num / per / 1e6, 4 * g.grid_num / per / 1e6)
    if name in ("k_mc_count", "k_mc_emit"):
        extra = "  %.0f GB/s of the 4 B/node field read" % (4 * g.grid_num / per / 1e6)
    print("  %-20s %8.4f ms/launch%s" % (name, per, extra))
