#!/usr/bin/env python
"""Per-kernel timing of the DFSPH step on the 1M dam-break scene via the library's own
CUDA-event profiler (wcsph_profile).  Usage: python tools/sweep_bench.py [steps] [nx ny nz]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from wcsph_b200 import _lib, dfsph, scenes  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dims = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (100, 100, 100)
pts, nl = scenes.dam_break(*dims)
dfsph.init_scene(pts, nl)
dfsph.reset_param()
for opt in ("list_build_v1",):
    if os.environ.get("WCSPH_OPT_" + opt.upper()):
        _lib.check(_lib.load().wcsph_set_option(dfsph.particle_data._ctx, opt.encode(), int(os.environ["WCSPH_OPT_" + opt.upper()])))
for _ in range(3):
    dfsph.step_fused(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    dfsph.step_fused(1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
L = _lib.load()
ctx = dfsph.particle_data._ctx
_lib.check(L.wcsph_profile(ctx, 1))
for _ in range(steps):
    dfsph.step_fused(1)
buf = C.create_string_buffer(1 << 16)
_lib.check(L.wcsph_profile_report(ctx, buf, len(buf)))
_lib.check(L.wcsph_profile(ctx, 0))
rows = []
for line in buf.value.decode().splitlines():
    n, c, t = line.split("\t")
    rows.append((float(t) / steps, int(c) / steps, n))
rows.sort(reverse=True)
print("lib %s  scene %s  NL %d  %.3f ms/step  %.1f M particle-steps/s  iters %s  flags %d" % (
    os.environ.get("WCSPH_LIB", "default"), dims, nl, ms, nl / ms / 1e3, (dfsph.vs_iter, dfsph.dv_iter, dfsph.pr_iter),
    dfsph.particle_data.hash_grid.status()))
tot = sum(r[0] for r in rows)
for t, c, n in rows:
    print("  %-46s %5.1f/step  %8.4f ms/step  %8.4f ms/launch  %5.1f%%" % (n, c, t, t / c, 100 * t / tot))
print("  sum of kernels %.3f ms/step" % tot)
