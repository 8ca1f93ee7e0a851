"""SURVEY 8(f) N2 -- marching-cubes reconstruction, CPU side: the case table compiled into the library against the
reference's MCData.txt (fixture tests/golden/mc_tables.npz), and the serial restatement of
MarchingCubeGrid.py:160-209,262-352 against properties a correct polygoniser must have."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tables(golden_dir):
    t = np.load(os.path.join(golden_dir, "mc_tables.npz"))
    return t["edgetable"], t["tritable"]


def test_packed_table_equals_reference_table(tables):
    """csrc/mc_table.h (one hex nibble per entry) decodes to MCData.txt's tritable; the edge table the reference also
    stores is the OR of the edges a case uses, so the library derives nothing else from the file."""
    edge, tri = tables
    src = open(os.path.join(ROOT, "wcsph_b200", "csrc", "mc_table.h")).read()
    hexs = "".join(re.findall(r'"([0-9a-f]+)"', src))
    assert len(hexs) == 256 * 16
    dec = np.array([-1 if ch == "f" else int(ch, 16) for ch in hexs], np.int32).reshape(256, 16)
    assert np.array_equal(dec, tri)
    derived = np.array([sum({1 << e for e in row if e >= 0}) for row in tri], np.int32)
    assert np.array_equal(derived, edge)
    assert int((tri >= 0).sum()) // 3 == 820                  # the classic table: 820 triangles over the 256 cases
    assert edge[0] == 0 and edge[255] == 0 and np.array_equal(edge, edge[::-1])      # complementary cases cut the same edges


def _sphere_oracle(n_side=14, radius=0.3):
    from oracle import oracle
    g = (np.arange(n_side) - (n_side - 1) / 2) * 0.05
    pts = np.array([[x, y, z] for x in g for y in g for z in g if x * x + y * y + z * z <= radius * radius], np.float32)
    box = np.array([[-0.6, -0.6, -0.6], [0.6, 0.6, 0.6]], np.float32)       # two far "solids" fix the bounding box
    allp = np.concatenate([pts, box]).astype(np.float32)
    m = oracle.McOracle(allp, len(pts))
    return m, allp, len(pts)


def test_oracle_blob_mesh_is_closed_and_near_the_blob(tables):
    edge, tri = tables
    m, allp, nl = _sphere_oracle()
    assert m.update_grid() == 0
    assert int(m.gridCount.sum()) == len(allp) and m.gridCount.max() <= 4
    rho = np.full(nl, 1000.0, np.float32)
    sv = m.cal_surface_point(rho)
    assert 0.75 < sv.max() < 0.9 and sv.min() == 0.0          # sum V_j W ~ 0.8 inside (V = 0.8 d^3, ParticleData.py:20), 0 far away
    n, v = m.marching_cube(edge, tri)
    assert n > 0 and n % 3 == 0 and len(v) == n
    # closed 2-manifold: every undirected edge is shared by exactly two triangles (vertex_interp orders its end points
    # -- MarchingCubeGrid.py:375-409 -- so that neighbouring cells produce bit-identical vertices)
    _, vid = np.unique(v, axis=0, return_inverse=True)
    t = vid.reshape(-1, 3)
    t = t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]      # drop degenerate slivers
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert np.all(counts % 2 == 0)
    r = np.linalg.norm(v, axis=1)
    assert 0.25 < r.min() and r.max() < 0.40                  # iso-surface hugs the blob of radius 0.3


def test_oracle_low_density_particles_and_vertex_cap(tables):
    edge, tri = tables
    m, allp, nl = _sphere_oracle(8, 0.2)
    m.update_grid()
    w0 = (8.0 / np.pi) / m.searchR ** 3
    low = np.full(nl, np.float32(m.liqiudMass * w0 * 0.999), np.float32)      # rho <= m W(0): skipped, :203
    assert not m.cal_surface_point(low).any()
    m.cal_surface_point(np.full(nl, 1000.0, np.float32))
    n, v = m.marching_cube(edge, tri)
    n2, v2 = m.marching_cube(edge, tri, max_vertex=300)
    assert n2 == n and len(v2) == 300 and np.array_equal(v2, v[:300])        # keeps counting, stops writing (:343-349)


def test_oracle_cell_overflow_keeps_first_four():
    from oracle import oracle
    p = np.array([[0.001 * k, 0.0, 0.0] for k in range(6)] + [[0.5, 0.5, 0.5], [-0.5, -0.5, -0.5]], np.float32)
    m = oracle.McOracle(p, 6)
    assert m.update_grid() == 2                               # "mc exceed grid" twice, :173-175
    c = int(np.argmax(m.gridCount))
    assert m.gridCount[c] == 4 and list(m.grid[c]) == [0, 1, 2, 3]


# ---- anisotropic branch (ParticleData.py:188-298, MarchingCubeGrid.py:215-243): restatement only, the GPU engine does not build it yet
@pytest.fixture(scope="module")
def aniso_scene():
    from oracle import oracle
    from wcsph_b200 import scenes
    pts, nl = scenes.dam_break(10, 10, 10, jitter=True, config_id=2)
    o = oracle.Oracle("dfsph", pts, nl, threads=8)
    o.call("update_grid")
    o.call("compute_density")
    mc = oracle.McOracle(pts, nl)
    pos_avr, G = oracle.cal_anistropic_kernel(o, mc.searchR)
    return o, mc, pts, nl, pos_avr, G


def test_oracle_anisotropy_matrix_matches_numpy_eigendecomposition(aniso_scene):
    """G_i = R diag(1 / (ks * [s0, max(s1, s0/kr), max(s2, s0/kr)])) R^T of the weighted covariance (ParticleData.py:243-279),
    recomputed here in float64 with numpy.linalg.eigh from the oracle's own neighbour table"""
    o, mc, pts, nl, pos_avr, G = aniso_scene
    pos = o.field("pos").astype(np.float64)
    nbr, cnt = o.field("neighbor"), o.field("neighborCount")
    R2 = 2.0 * np.float32(mc.searchR)
    rng = np.random.default_rng(3)
    checked = 0
    for i in rng.choice(nl, 60, replace=False):
        js = nbr[i, :cnt[i]]
        js = js[js < nl]
        d = np.linalg.norm(pos[i] - pos[js], axis=1)
        w = np.where(d < R2, 1.0 - (d / R2) ** 3, 0.0)
        assert np.allclose(pos_avr[i], (w[:, None] * pos[js]).sum(0) / w.sum(), atol=2e-6)
        if cnt[i] <= 25:
            assert np.array_equal(G[i], 0.5 * np.eye(3, dtype=np.float32))
            continue
        r = pos[js] - pos_avr[i].astype(np.float64)
        Cm = (w[:, None, None] * r[:, :, None] * r[:, None, :]).sum(0) / w.sum()
        s, Rm = np.linalg.eigh(Cm)
        s, Rm = s[::-1], Rm[:, ::-1]
        inv = 1.0 / (1400.0 * np.array([s[0], max(s[1], s[0] / 4.0), max(s[2], s[0] / 4.0)]))
        assert np.allclose(G[i], Rm @ np.diag(inv) @ Rm.T, rtol=2e-3, atol=2e-4 * inv.max())
        checked += 1
    assert checked >= 40
    ev = np.linalg.eigvalsh(G.astype(np.float64))
    assert np.all(ev > 0) and np.all(ev[:, 2] / ev[:, 0] <= 4.0 * (1 + 1e-4))          # kr bounds the stretch
    assert np.abs(G - G.transpose(0, 2, 1)).max() < 1e-6


def test_oracle_color_map_and_anisotropic_surface(aniso_scene, tables):
    from oracle import oracle
    edge, tri = tables
    o, mc, pts, nl, pos_avr, G = aniso_scene
    color, grad = oracle.compute_color_map(o)
    pos = o.field("pos")[:nl]
    lo, hi = pos.min(0), pos.max(0)
    inner = np.all((pos > lo + 0.12) & (pos < hi - 0.12), axis=1)
    assert inner.sum() > 50
    assert np.all((color[inner] > 0.6) & (color[inner] < 1.3))                          # sum_j m/rho_j W ~ 1 in the bulk
    g = np.linalg.norm(grad, axis=1)
    assert g[inner].mean() < 0.3 * g[~inner].mean()                                     # the colour gradient lives at the surface
    mc.update_grid(o.field("pos"))
    sv = mc.cal_surface_point_anistropic(o.field("rho"), pos_avr, G)
    assert sv.max() > 0.5 and sv.min() == 0.0
    n, v = mc.marching_cube(edge, tri)
    assert n > 0 and n % 3 == 0
    assert np.all(v.min(0) > lo - 0.1) and np.all(v.max(0) < hi + 0.1)
