"""SURVEY 8(f) N1 -- canvas parity: the CUDA splat pass (C ABI wcsph_canvas_*) against the serial CPU
restatement of Canvas.py:138-209 + the scripts' draw_particle kernels.  Integer / pixel work: bit-exact
(image and depth buffer compared with array_equal)."""
import numpy as np
import pytest

from .util import make_engine, scene

pytestmark = pytest.mark.gpu


def _oracle_picture(m, nl, style):
    from oracle import oracle
    c = m.sph_canvas
    pos = m.particle_data.pos.to_numpy()
    return oracle.canvas_draw_particle(pos, nl, c.view[0], c.proj[0], c.sizex, c.sizey, style)


def _check(m, nl, style, min_lit):
    c = m.sph_canvas
    c.clear_canvas()
    m.draw_particle()
    img, depth = c.img.to_numpy(), c.depth.to_numpy()
    oi, od = _oracle_picture(m, nl, style)
    assert img.shape == (c.sizex, c.sizey, 3) and depth.shape == (c.sizex, c.sizey)
    assert np.count_nonzero(oi[:, :, 0]) >= min_lit, "degenerate test picture"
    assert np.array_equal(img, oi), "image differs in %d pixels" % int(np.count_nonzero((img != oi).any(axis=2)))
    assert np.array_equal(depth, od), "depth differs in %d pixels" % int(np.count_nonzero(depth != od))


@pytest.mark.parametrize("solver,style", [("dfsph", 1), ("sesph", 0), ("iisph", 0), ("pcisph", 0)])
def test_canvas_static_cam_matches_restatement(solver, style):
    """the as-shipped scene of every script, drawn the way its main loop does (dfsph.py:604,621-623)."""
    pts, nl = scene(solver)
    m = make_engine(solver, pts, nl)
    target = (0.0, 1.0, 0.0) if solver in ("dfsph", "iisph") else (0.0, 0.0, 0.0)
    m.sph_canvas.static_cam(*target)
    _check(m, nl, style, 1000)
    for _ in range(5):
        m.step()
    _check(m, nl, style, 1000)


def test_canvas_perspective_cams():
    """yaw_cam / pitch_cam (Canvas.py:37-54): perspective divide, particles crossing pixel borders while the camera turns."""
    pts, nl = scene("dfsph", "dam32")
    m = make_engine("dfsph", pts, nl)
    m.step_fused(3)
    c = m.sph_canvas
    c.set_target(0.8, 0.8, 0.8)
    for k in range(40):
        c.yaw_cam(0.8, 0.8, 0.8)
        if k % 13 == 0:
            _check(m, nl, 1, 1000)
    c.yaw = 0.0
    for k in range(30):
        c.pitch_cam(0.8, 0.8, 0.8)
    _check(m, nl, 1, 1000)
    c.set_view_point(2.5, 0.4, 0.0, 1.2)         # eye inside the tank: particles behind the eye and beyond the borders
    _check(m, nl, 1, 100)


def test_canvas_1m_and_png(tmp_path):
    """BASELINE configs[1] (1M liquid + 131,808 boundary): heavy depth complexity (4 particles per pixel column)."""
    from wcsph_b200 import scenes
    pts, nl = scenes.dam_break(100, 100, 100, jitter=True, config_id=2)
    m = make_engine("dfsph", pts, nl)
    m.step_fused(2)
    m.sph_canvas.static_cam(2.5, 2.5, 0.0)
    m.sph_canvas.set_fov(2.6)
    _check(m, nl, 1, 20000)
    import time
    t0 = time.perf_counter()
    _oracle_picture(m, nl, 1)
    print("\n[canvas 1M] serial CPU restatement: %.1f ms per frame" % ((time.perf_counter() - t0) * 1e3))
    path = tmp_path / "frame.png"
    m.sph_canvas.write_png(str(path))
    data = path.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and len(data) > 1000
    # device views stay on the GPU (a consumer that never leaves HBM)
    t = m.sph_canvas.img.to_torch()
    assert t.is_cuda and tuple(t.shape) == (512, 512, 3)


def test_reference_main_loop_with_canvas_and_surface(tmp_path, monkeypatch, capsys):
    """dfsph.py:595-646 as a function: step, draw, console line, PNG frames, marching-cubes OBJ -- nothing but files and stdout."""
    import importlib
    monkeypatch.chdir(tmp_path)
    m = importlib.reload(importlib.import_module("wcsph_b200.dfsph"))
    m.main(steps=3, png_every=1, surface=True)
    out = capsys.readouterr().out
    assert out.count("viscorcity:") == 3 and "time:" in out
    assert (tmp_path / "1.png").exists() and (tmp_path / "3.png").exists()
    assert (tmp_path / "out" / "mc_0.obj").exists()
    assert m.particle_data.mc_grid.frame == 1 and m.particle_data.hash_grid.status() == 0
