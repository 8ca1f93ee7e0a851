"""SURVEY 8(f) N1 -- the canvas: host camera math, PNG encoding and the CPU restatement of
Canvas.py:138-209 against closed-form expectations (the reference holds no golden image: parity unpinned)."""
import struct
import zlib

import numpy as np
import pytest

from wcsph_b200.Canvas import Canvas, encode_png

BRESENHAM_R3 = {(0, 3), (0, -3), (3, 0), (-3, 0), (1, 3), (-1, 3), (1, -3), (-1, -3),
                (3, 1), (-3, 1), (3, -1), (-3, -1), (2, 2), (-2, 2), (2, -2), (-2, -2)}


def test_static_cam_matrices():
    """Canvas.py:57-63,67-100: eye = target + (0,0,3), 'ortho' projection with fov 2."""
    c = Canvas(512, 512)
    c.static_cam(0.0, 1.0, 0.0)
    assert np.allclose(c.eye, [0.0, 1.0, 3.0]) and c.ortho == 1 and c.fov == 2.0
    view, proj = c.view.to_numpy()[0], c.proj.to_numpy()[0]
    assert view.dtype == np.float32 and view.shape == (4, 4)
    assert np.allclose(view, [[1, 0, 0, 0], [0, 1, 0, -1], [0, 0, 1, -3], [0, 0, 0, 1]])
    s = 1.0 / np.tan(1.0)
    assert np.allclose(proj, [[s, 0, 0, 0], [0, s, 0, 0], [0, 0, 1 / (1 - 1000.0), 1 / (1 - 1000.0)], [0, 0, 0, 1]])


def test_yaw_and_pitch_cam_step():
    """Canvas.py:37-54: 0.003 rad per call, perspective projection, distance 3."""
    c = Canvas(640, 480)
    for _ in range(10):
        c.yaw_cam(0.0, 1.0, 0.0)
    assert c.yaw == pytest.approx(0.03) and c.ortho == 0 and c.fov == 1.0
    assert np.linalg.norm(c.eye - c.target) == pytest.approx(3.0)
    proj = c.proj.to_numpy()[0]
    ys = 1.0 / np.tan(0.5)
    assert proj[1, 1] == pytest.approx(ys) and proj[0, 0] == pytest.approx(ys / (640 / 480)) and proj[3, 2] == -1.0
    view = c.view.to_numpy()[0]
    assert np.allclose(view[:3, :3] @ view[:3, :3].T, np.eye(3), atol=1e-6)         # orthonormal camera axes
    assert np.allclose(view @ np.append(c.eye, 1.0), [0, 0, 0, 1], atol=1e-6)       # the eye is the origin
    p = Canvas(64, 64)
    p.pitch = 5.0
    p.update_cam()
    assert p.pitch == 1.57                                                          # clamp, Canvas.py:69-70
    q = Canvas(64, 64)
    for _ in range(200):
        q.pitch_cam(0, 0, 0)
    assert 0.5 <= q.pitch < 0.504                                                   # stops stepping at 0.5


def _draw(pos, nl, style, canvas):
    from oracle import oracle
    return oracle.canvas_draw_particle(pos, nl, canvas.view[0], canvas.proj[0], canvas.sizex, canvas.sizey, style)


def test_oracle_single_liquid_is_a_radius3_outline():
    c = Canvas(512, 512)
    c.static_cam(0.0, 1.0, 0.0)
    img, depth = _draw(np.array([[0.0, 1.0, 0.0]], np.float32), 1, 0, c)
    lit = {(int(x) - 256, int(y) - 256) for x, y in zip(*np.nonzero(img[:, :, 0]))}
    assert lit == BRESENHAM_R3
    assert np.all(img[img[:, :, 0] > 0] == 1.0)
    zs = depth[img[:, :, 0] > 0]
    assert np.all(zs == zs[0]) and 0.0 < zs[0] < 1.0
    assert np.all(depth[img[:, :, 0] == 0] == 1.0)                                  # clear value, Canvas.py:209
    img1, _ = _draw(np.array([[0.0, 1.0, 0.0]], np.float32), 1, 1, c)               # dfsph.py:591-593: + grey centre
    assert img1[256, 256, 0] == pytest.approx(0.3) and np.count_nonzero(img1[:, :, 0]) == 17


def test_oracle_depth_test_and_tie_order():
    c = Canvas(128, 128)
    c.static_cam(0.0, 0.0, 0.0)
    # a solid in front of (larger z = nearer the eye at +z) / behind a liquid's outline pixel (0,+3)
    px = 2.0 / (c.proj[0][0, 0] * 128)                                              # world units per pixel
    liquid = [0.0, 0.0, 0.0]
    front, back = [0.0, 3.2 * px, 0.5], [0.0, 3.2 * px, -0.5]
    img, _ = _draw(np.array([liquid, front], np.float32), 1, 0, c)
    assert img[64, 67, 0] == pytest.approx(0.3)
    img, _ = _draw(np.array([liquid, back], np.float32), 1, 0, c)
    assert img[64, 67, 0] == 1.0
    same = [0.0, 3.2 * px, 0.0]                                                     # equal depth: the earlier fragment stays
    img, _ = _draw(np.array([liquid, same], np.float32), 1, 0, c)
    assert img[64, 67, 0] == 1.0
    # off-screen and behind-the-far-plane particles leave the canvas untouched
    img, depth = _draw(np.array([[100.0, 0.0, 0.0], [0.0, 0.0, -2000.0]], np.float32), 2, 0, c)
    assert not img.any() and np.all(depth == 1.0)


def test_png_encoding_roundtrip():
    img = np.zeros((4, 3, 3), np.float32)          # sx = 4, sy = 3
    img[0, 0] = (1.0, 0.0, 0.0)                    # bottom-left in canvas coordinates (y up)
    img[3, 2] = (0.0, 0.0, 2.0)                    # top-right, clipped to 1
    data = encode_png(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, bits, ctype = struct.unpack(">IIBB", data[16:26])
    assert (w, h, bits, ctype) == (4, 3, 8, 2)
    n = struct.unpack(">I", data[33:37])[0]
    assert data[37:41] == b"IDAT"
    raw = zlib.decompress(data[41:41 + n])
    rows = np.frombuffer(raw, np.uint8).reshape(3, 1 + 4 * 3)
    assert np.all(rows[:, 0] == 0)
    px = rows[:, 1:].reshape(3, 4, 3)
    assert tuple(px[2, 0]) == (255, 0, 0) and tuple(px[0, 3]) == (0, 0, 255)
