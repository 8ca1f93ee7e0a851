"""Canvas -- the particle-splat view the reference's step loops draw into.

Mirror of the reference's `Canvas(sizex, sizey)` (Canvas.py:7-209): the same camera state and camera
methods (`static_cam`, `yaw_cam`, `pitch_cam`, `set_view_point`, `set_fov`, `set_target`, `update_cam`),
`view` / `proj` / `img` / `depth` with `.to_numpy()`, `clear_canvas()`, `export_png*`.  The per-particle
device functions (`draw_sphere`, `draw_point`, Canvas.py:150-201) cannot be called one particle at a time
from Python; what the scripts do with them -- their `draw_particle` kernels -- is `draw_particle(pd, style)`
here, one launch over the cell-sorted device positions (`csrc/canvas.cu`).  Pixels are 64-bit
(depth, colour) keys resolved with atomicMin, so the picture is the reference's *serial* result and is
reproducible, which the reference's own racy depth test is not.

No CPU fallback: a missing library raises.
"""
import ctypes as C
import math
import struct
import zlib

import numpy as np

from . import _lib

STYLE_SPLIT = 0     # sesph.py:201-207, pcisph.py:288-293, iisph.py:401-406
STYLE_DFSPH = 1     # dfsph.py:585-593 (a grey point for the liquids too)


def _unit(v):
    return v / np.linalg.norm(v)


def _look_at(eye, target, up):
    """rows = camera axes, last column = -axis . eye (Canvas.py:77-90)."""
    back = _unit(eye - target)
    right = _unit(np.cross(up, back))
    above = np.cross(back, right)
    m = np.eye(4)
    for row, axis in enumerate((right, above, back)):
        m[row, :3] = axis
        m[row, 3] = -np.dot(axis, eye)
    return m


def _projection(fov, ratio, near, far, ortho):
    """Canvas.py:82-83,93-99: perspective (w = -z) or the reference's 'ortho' variant (w = 1)."""
    ys = 1.0 / math.tan(fov / 2.0)
    xs = ys / ratio
    span = near - far
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = xs, ys
    if ortho:
        m[2, 2], m[2, 3], m[3, 3] = 1.0 / span, near / span, 1.0
    else:
        m[2, 2], m[2, 3], m[3, 2] = far / span, near * far / span, -1.0
    return m


def encode_png(img):
    """img[sx, sy, 3] f32 in [0, 1] (x to the right, y up, like ti.imwrite) -> 8-bit RGB PNG bytes."""
    rgb = (np.clip(np.asarray(img, np.float32), 0.0, 1.0) * 255.0).astype(np.uint8).transpose(1, 0, 2)[::-1]
    h, w = rgb.shape[:2]
    raw = b"".join(b"\x00" + rgb[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


class _Mat44Field:
    """stand-in for ti.Matrix.field(4, 4, f32, shape=(1)): to_numpy() / from_numpy() / [0]."""

    def __init__(self):
        self._a = np.zeros((1, 4, 4), np.float32)

    def to_numpy(self):
        return self._a.copy()

    def from_numpy(self, a):
        self._a[...] = np.asarray(a, np.float32).reshape(1, 4, 4)

    def __getitem__(self, i):
        return self._a[i]


class _PixelField:
    def __init__(self, canvas, which):
        self._canvas, self._which = canvas, which

    @property
    def shape(self):
        c = self._canvas
        return (c.sizex, c.sizey, 3) if self._which == 0 else (c.sizex, c.sizey)

    def to_numpy(self):
        return self._canvas._resolve()[self._which]

    def to_torch(self):
        """device views (no copy): img [sx, sy, 3] or depth [sx, sy], valid until the next resolve."""
        self._canvas._resolve(host=False)
        return self._canvas._img_t if self._which == 0 else self._canvas._depth_t


class Canvas:
    def __init__(self, sizex, sizey):
        self.sizex, self.sizey = int(sizex), int(sizey)
        self.screenRes = np.array([self.sizex, self.sizey])
        self.view, self.proj = _Mat44Field(), _Mat44Field()
        self.img, self.depth = _PixelField(self, 0), _PixelField(self, 1)
        # camera defaults, Canvas.py:16-33
        self.eye = np.array([0.0, 0.0, 1.0])
        self.target = np.array([0.0, 0.0, 0.0])
        self.up = np.array([0.0, 1.0, 0.0])
        self.ratio = sizex / sizey
        self.yaw = self.pitch = self.roll = 0.0
        self.scale = 1.0
        self.fov, self.near, self.far, self.ortho = 1.0, 1.0, 1000.0, 0
        self.frame, self.fps = 0, 30.0
        self._zbuf = self._img_t = self._depth_t = None
        self._pd = None

    # ---- camera (host side, float64 like the reference, stored as f32) ----
    def update_cam(self):
        """Canvas.py:67-100."""
        self.pitch = max(min(self.pitch, 1.57), -1.57)
        cp, sp, cy, sy = math.cos(self.pitch), math.sin(self.pitch), math.cos(self.yaw), math.sin(self.yaw)
        self.eye[:] = self.target + np.array([self.scale * cp * sy, self.scale * sp, self.scale * cp * cy])
        self.up[:] = (-sp * sy, cp, -sp * cy)
        self.view.from_numpy(_look_at(self.eye, self.target, self.up))
        self.proj.from_numpy(_projection(self.fov, self.ratio, self.near, self.far, self.ortho))

    def set_view_point(self, yaw, pitch, roll, scale):
        self.yaw, self.pitch, self.roll, self.scale = yaw, pitch, roll, scale
        self.update_cam()

    def set_fov(self, fov):
        self.fov = fov
        self.update_cam()

    def set_target(self, targetx, targety, targetz):
        self.target[:] = (targetx, targety, targetz)
        self.update_cam()

    def _aim(self, fov, ortho, t):
        self.fov, self.ortho = fov, ortho
        self.target[:] = t

    def static_cam(self, targetx, targety, targetz):
        """Canvas.py:57-63."""
        self._aim(2.0, 1, (targetx, targety, targetz))
        self.set_view_point(0.0, 0.0, 0.0, 3.0)

    def yaw_cam(self, targetx, targety, targetz):
        """Canvas.py:37-44: one 0.003 rad step per call until yaw reaches 3.14."""
        self._aim(1.0, 0, (targetx, targety, targetz))
        if self.yaw < 3.14:
            self.set_view_point(self.yaw + 0.003, 0.0, 0.0, 3.0)

    def pitch_cam(self, targetx, targety, targetz):
        """Canvas.py:47-54."""
        self._aim(1.0, 0, (targetx, targety, targetz))
        if self.pitch < 0.5:
            self.set_view_point(0.0, self.pitch + 0.003, 0.0, 3.0)

    # ---- device side ----
    def bind(self, particle_data):
        """the ParticleData whose context / stream the canvas draws with (the solver modules call this)."""
        self._pd = particle_data
        return self

    def _buffers(self, pd):
        import torch
        if pd is None or pd._ctx is None:
            raise _lib.WcsphError("Canvas: the ParticleData has no device context yet (setup_data_gpu first)")
        if self._zbuf is None:
            n = self.sizex * self.sizey
            self._zbuf = torch.empty(n, dtype=torch.int64, device="cuda")
            self._img_t = torch.empty((self.sizex, self.sizey, 3), dtype=torch.float32, device="cuda")
            self._depth_t = torch.empty((self.sizex, self.sizey), dtype=torch.float32, device="cuda")
        self._pd = pd
        return C.c_void_p(self._zbuf.data_ptr())

    def clear_canvas(self, particle_data=None):
        """Canvas.py:205-209.  The first call needs the ParticleData whose stream / context the canvas rides on."""
        pd = particle_data if particle_data is not None else self._pd
        z = self._buffers(pd)
        _lib.check(_lib.load().wcsph_canvas_clear(pd._ctx, z, self.sizex, self.sizey))

    def draw_particle(self, particle_data, style=STYLE_SPLIT):
        """the scripts' draw_particle kernels (dfsph.py:585-593 style 1; sesph.py:201-207 style 0)."""
        z = self._buffers(particle_data)
        v = np.ascontiguousarray(self.view[0], np.float32)
        p = np.ascontiguousarray(self.proj[0], np.float32)
        _lib.check(_lib.load().wcsph_canvas_draw_particle(particle_data._ctx, v.ctypes.data, p.ctypes.data,
                                                          self.sizex, self.sizey, int(style), z))
        if getattr(particle_data, "world_size", 1) > 1:
            # z-slab ranks: per pixel the smallest key of any rank (keys are < 2^34, int64 order = key order)
            import torch.distributed as dist
            particle_data.sync()
            dist.all_reduce(self._zbuf, op=dist.ReduceOp.MIN)

    def _resolve(self, host=True):
        if self._zbuf is None:
            raise _lib.WcsphError("Canvas: nothing drawn yet (clear_canvas / draw_particle first)")
        pd = self._pd
        _lib.check(_lib.load().wcsph_canvas_resolve(pd._ctx, C.c_void_p(self._zbuf.data_ptr()), self.sizex, self.sizey,
                                                    C.c_void_p(self._img_t.data_ptr()), C.c_void_p(self._depth_t.data_ptr())))
        pd.sync()
        if host:
            return self._img_t.cpu().numpy(), self._depth_t.cpu().numpy()
        return None

    # ---- PNG (ti.imwrite, Canvas.py:123-134) ----
    def write_png(self, path):
        with open(path, "wb") as f:
            f.write(encode_png(self.img.to_numpy()))

    def export_png(self, time):
        """Canvas.py:123-127: one frame per 1/fps of simulated time."""
        if int(time * self.fps) == self.frame:
            self.write_png(str(self.frame) + ".png")
            self.frame += 1

    def export_png_frame(self, frame):
        """Canvas.py:130-134."""
        if frame > self.frame:
            self.frame = frame
            self.write_png(str(self.frame) + ".png")
