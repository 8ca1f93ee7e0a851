"""Pins the CPU oracle against everything the reference lets us evaluate without Taichi
(SURVEY.md section 4 / 8c).  The reference ships no tests or golden vectors, so these anchors
are: its own data files, its pure-numpy host code, closed-form kernel identities, and the
t=0 hash statistics derived independently in SURVEY.md."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as orc
from wcsph_b200 import scenes


@pytest.fixture(scope="module")
def anchors(golden_dir):
    return json.load(open(os.path.join(golden_dir, "anchors.json")))


def test_liquid_block_matches_reference_dump(golden_dir):
    # model/liqiud.obj is the reference's own dump of dfsph.py:70-73 ("%f": 6 decimals)
    liq = np.load(os.path.join(golden_dir, "liqiud.npy"))
    mine = scenes.dfsph_liquid_block()
    assert liq.shape == mine.shape == (8000, 3)
    assert np.max(np.abs(liq - mine)) < 1e-6


def test_box_boundary_fixture(golden_dir, anchors):
    box = np.load(os.path.join(golden_dir, "box_boundry.npy"))
    assert box.shape == (anchors["box_boundry_count"], 3) == (25387, 3)
    assert box.min(0).tolist() == [-1.0, 0.0, -1.0] and box.max(0).tolist() == [1.0, 2.0, 1.0]


def test_pci_coff_matches_reference_source(anchors):
    # anchors.json value = pcisph.py:74-115 executed verbatim by tests/golden/make_fixtures.py
    assert anchors["pci_coff"] == pytest.approx(0.004597319327225708, rel=1e-15)
    assert orc.get_pci_coff() == pytest.approx(anchors["pci_coff"], rel=1e-12)


@pytest.mark.parametrize("solver,w0", [("dfsph", 2546.479089), ("sesph", 2546.479133)])
def test_cubic_kernel_identities(solver, w0):
    L = orc.lib()
    c = orc.solver_constants(solver)
    p = orc.make_params(c)
    style = c["style"]
    h = c["searchR"]
    W = lambda r: L.oracle_cubic_W_norm(C.byref(p), r, style)
    assert W(0.0) == pytest.approx(w0, rel=2e-7)                       # 8/(pi h^3), SURVEY 2.3
    assert W(0.0) == pytest.approx(8.0 / (c["pi"] * h ** 3), rel=2e-7)
    assert W(h / 2) == pytest.approx(0.25 * W(0.0), rel=1e-6)
    assert W(h * 1.0001) == 0.0 and W(2 * h) == 0.0
    g = np.zeros(3, dtype=np.float32)
    r0 = np.zeros(3, dtype=np.float32)
    L.oracle_cubic_gradW(C.byref(p), r0.ctypes.data, g.ctypes.data, style)
    assert np.all(g == 0.0)                                           # guard rl > 1e-5
    # gradient is the derivative of W
    r = np.array([0.03, 0.0, 0.0], dtype=np.float32)
    L.oracle_cubic_gradW(C.byref(p), r.ctypes.data, g.ctypes.data, style)
    d = 1e-4
    assert g[0] == pytest.approx((W(0.03 + d) - W(0.03 - d)) / (2 * d), rel=2e-3)
    # unit integral: sum over a fine lattice
    n = 40
    xs = (np.arange(-n, n + 1) + 0.5) * (h / n)
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")
    R = np.sqrt(X * X + Y * Y + Z * Z).ravel()
    q = R / h
    Wv = np.where(q <= 0.5, 6 * q ** 3 - 6 * q ** 2 + 1, np.where(q <= 1, 2 * (1 - q) ** 3, 0.0)) * 8.0 / (c["pi"] * h ** 3)
    assert Wv.sum() * (h / n) ** 3 == pytest.approx(1.0, abs=2e-3)


def test_cohesion_adhesion_kernels():
    L = orc.lib()
    p = orc.make_params(orc.solver_constants("dfsph"))
    h = 0.1
    assert L.oracle_cohesion_W_norm(C.byref(p), 0.11) == 0.0
    r = 0.07
    assert L.oracle_cohesion_W_norm(C.byref(p), r) == pytest.approx(32 / (math.pi * h ** 9) * (h - r) ** 3 * r ** 3, rel=1e-4)
    r = 0.03
    # as written in CohesionKernel.py:27 the h^6/64 term is NOT scaled by m_k (replicated)
    assert L.oracle_cohesion_W_norm(C.byref(p), r) == pytest.approx(
        32 / (math.pi * h ** 9) * 2 * (h - r) ** 3 * r ** 3 - h ** 6 / 64, rel=1e-4)
    assert L.oracle_adhesion_W_norm(C.byref(p), 0.04) == 0.0
    r = 0.075
    assert L.oracle_adhesion_W_norm(C.byref(p), r) == pytest.approx(
        0.007 / h ** 3.25 * (-4 * r * r / h + 6 * r - 2 * h) ** 0.25, rel=1e-4)


def test_hash_statistics_dfsph_scene():
    """N 33,387, grid 41^3, max bucket occupancy 13, candidates mean
    227.297 max 312, fraction with an in-range double count 0.084375."""
    pts, nl = scenes.scene_dfsph()
    o = orc.Oracle("dfsph", pts, nl, threads=8)
    o.call("update_grid")
    assert o.count == 33387 and nl == 8000
    assert o.field("blockSize").tolist() == [41, 41, 41]
    assert int(o.field("gridCount").max()) == 13
    nc = o.field("neighborCount")
    assert nc.mean() == pytest.approx(227.297, abs=5e-4) and int(nc.max()) == 312 and int(nc.min()) >= 125
    assert o.flag("exceed_grid") == 0 and o.flag("exceed_neighbor") == 0
    nb, pos = o.field("neighbor"), o.field("pos")
    dup = 0
    for i in range(nl):
        js = nb[i, :nc[i]]
        d = np.linalg.norm(pos[js] - pos[i], axis=1)      # in range: |r| <= searchR (lattice pairs at exactly h count)
        jj = js[d <= 0.1]
        dup += len(np.unique(jj)) < len(jj)
    assert dup / nl == pytest.approx(0.084375, abs=1e-6)


def test_hash_statistics_sesph_scene():
    """9,128 solids, N 17,128, grid 21^3, max occupancy 28, candidates
    mean 839.206 max 1184."""
    pts, nl = scenes.scene_sesph()
    o = orc.Oracle("sesph", pts, nl, threads=8)
    o.call("update_grid")
    assert o.count == 17128 and o.count - nl == 9128
    assert o.field("blockSize").tolist() == [21, 21, 21]
    assert int(o.field("gridCount").max()) == 28
    nc = o.field("neighborCount")
    assert nc.mean() == pytest.approx(839.206, abs=5e-4) and int(nc.max()) == 1184


def test_dam_break_generator_counts():
    # NS = bx*by*bz - (bx-2)(by-2)(bz-2)
    pts, nl = scenes.dam_break(10, 10, 10)
    bx, by, bz = 22, 17, 12
    assert nl == 1000 and len(pts) - nl == bx * by * bz - (bx - 2) * (by - 2) * (bz - 2)


@pytest.mark.parametrize("solver", ["sesph", "dfsph", "iisph", "pcisph"])
def test_oracle_runs_and_stays_finite(solver):
    pts, nl = getattr(scenes, "scene_" + solver)()
    o = orc.Oracle(solver, pts, nl, threads=8)
    for _ in range(3):
        o.step()
    assert np.all(np.isfinite(o.field("pos")))
    assert np.all(np.isfinite(o.field("vel")))
    if solver == "dfsph":
        assert (o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter")) == (1, 1, 2)
    if solver == "pcisph":
        assert o.flag("pr_iter") == 3


@pytest.mark.parametrize("solver", ["dfsph", "sesph", "iisph", "pcisph"])
def test_oracle_reproduces_committed_goldens(solver, golden_dir):
    """tests/golden/oracle_steps.npz (made by make_fixtures.py oracle): the restatement is deterministic at
    one thread, so a change in it shows up here before it silently moves the parity target"""
    z = np.load(os.path.join(golden_dir, "oracle_steps.npz"))
    pts, nl = getattr(scenes, "scene_" + solver)()
    o = orc.Oracle(solver, pts, nl, threads=1)
    for _ in range(int(z[solver + "_iters"][3])):
        o.step()
    assert (o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter")) == tuple(int(x) for x in z[solver + "_iters"][:3])
    assert np.array_equal(o.field("rho"), z[solver + "_rho"])
    assert np.array_equal(o.field("pos")[:nl], z[solver + "_pos"])
    assert np.array_equal(o.field("vel"), z[solver + "_vel"])


def test_oracle_reproduces_canvas_and_mc_goldens(golden_dir):
    """committed canvas / marching-cubes goldens (tests/golden/canvas_mc_goldens.json, made by make_fixtures.py canvas_mc):
    the restatement is deterministic and has not drifted"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(golden_dir, "make_fixtures.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    mf.REF = None                                   # nothing here may read /root/reference: the tables come from the fixture
    t = np.load(os.path.join(golden_dir, "mc_tables.npz"))
    mf.mc_tables = lambda: (t["edgetable"], t["tritable"])
    want = json.load(open(os.path.join(golden_dir, "canvas_mc_goldens.json")))
    got = mf.canvas_mc_goldens()
    assert got == want
