#!/usr/bin/env python
"""Timing of the canvas pass on the 1M dam-break scene: clear + draw_particle + resolve per
frame, per kernel via the library's CUDA-event profiler, and end to end including the D2H of the image
(what `gui.set_image(sph_canvas.img.to_numpy())`, dfsph.py:623, costs).  The serial CPU restatement is timed by
tests/test_canvas_gpu.py::test_canvas_1m_and_png (-s prints it): only tests/ and bench.py may execute oracle/.  Usage: python tools/canvas_bench.py [frames]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from wcsph_b200 import _lib, dfsph, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pts, nl = scenes.dam_break(100, 100, 100, jitter=True, config_id=2)
dfsph.init_scene(pts, nl)
dfsph.reset_param()
dfsph.step_fused(3)
cv = dfsph.sph_canvas
cv.static_cam(2.5, 2.5, 0.0)
cv.set_fov(2.6)


def frame():
    cv.clear_canvas()
    dfsph.draw_particle()


for _ in range(3):
    frame()
    cv.img.to_torch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(frames):
    frame()
    cv.img.to_torch()
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / frames
t0 = time.perf_counter()
for _ in range(frames):
    frame()
    img = cv.img.to_numpy()
e2e_ms = (time.perf_counter() - t0) / frames * 1e3
L = _lib.load()
ctx = dfsph.particle_data._ctx
_lib.check(L.wcsph_profile(ctx, 1))
for _ in range(frames):
    frame()
    cv.img.to_torch()
buf = C.create_string_buffer(1 << 16)
_lib.check(L.wcsph_profile_report(ctx, buf, len(buf)))
_lib.check(L.wcsph_profile(ctx, 0))
n_total = len(pts)
print("canvas 512x512, %d particles (%d liquid): %.3f ms/frame on the device, %.3f ms/frame incl. D2H of img (3 MB)" % (
    n_total, nl, dev_ms, e2e_ms))
for line in buf.value.decode().splitlines():
    n, c, t = line.split("\t")
    per = float(t) / int(c)
    extra = ""
    if n == "k_canvas_draw":
        extra = "  %.1f G particles/s, %.0f GB/s of the 16 B/particle position stream" % (n_total / per / 1e6, 16 * n_total / per / 1e6)
    print("  %-20s %8.4f ms/launch%s" % (n, per, extra))
