"""SURVEY 8(f) N3 on the GPU: csrc/boundry.cu through wcsph_b200.boundry against (1) the executed reference, (2) the restatement."""
import os

import numpy as np
import pytest

from tests.test_boundry_cpu import load_golden

pytestmark = pytest.mark.gpu


def _write_obj(path, verts, faces):
    with open(path, "w") as fo:
        for p in verts:
            fo.write("v %.6f %.6f %.6f\n" % tuple(p))
        for t in faces:
            fo.write("f %d %d %d\n" % tuple(t))


def test_boundry_cuda_matches_reference_executed(tmp_path):
    """same mesh, the initial point set the reference drew injected: sorted arrays, hash map, phase groups, the samples and
    their order -- all bit-exact against what the unmodified boundry.py computed."""
    from wcsph_b200 import boundry as bd
    z, meta = load_golden()
    n = meta["numInitialPoints"]
    obj = os.path.join(tmp_path, "box.obj")
    _write_obj(obj, meta["mesh"]["vertices"], meta["mesh"]["faces"])
    assert bd.loadObj(obj) == n
    assert (bd.padding_num, bd.hash_map_size, bd.phase_vec_max, bd.faceNum) == (meta["padding_num"], meta["hash_map_size"], meta["phase_vec_max"], meta["faceNum"])
    assert np.array_equal(bd._s["tri_normal"], z["tri_normal"]) and np.array_equal(bd._s["tri_area"], z["tri_area"])
    bd.set_initial_points(z["init_pos"][:n], z["init_id"][:n])
    assert np.array_equal(bd.fetch("cell")[:, :3], z["init_cell"])
    bd.gpu_bitonic_sort()
    assert np.array_equal(bd.fetch("cell")[:, :3], z["sorted_cell"])
    pos = bd.fetch("pos")
    assert np.array_equal(pos[:, :3], z["sorted_pos"]) and np.array_equal(pos[:, 3].view(np.int32), z["sorted_id"])
    bd.build_hmap()
    assert np.array_equal(bd.fetch("start_index"), z["hmap_start_index"]) and np.array_equal(bd.fetch("hcell")[:, :3], z["hmap_cell"])
    assert np.array_equal(bd.fetch("phase_group_count"), z["phase_group_count"])
    assert np.array_equal(bd.fetch("phase_group")[:, :, :3], z["phase_group"])
    assert int(bd.fetch("counters")[1]) == int(z["hash_count"][0]) and bd.detect_hmap()
    counts = []
    for pg, trial in bd.launch_order():
        bd.possion_disk_sample(pg, trial)
        counts.append(int(bd.fetch("counters")[0]))
    assert np.array_equal(np.array(counts), z["launch"])
    ns = counts[-1]
    assert ns == meta["sample_count"]
    assert np.array_equal(bd.fetch("possion_sample")[:ns], z["possion_sample"])
    assert np.array_equal(bd.fetch("sample_count"), z["hmap_sample_count"])


def test_boundry_cuda_own_points_match_oracle_and_are_poisson_disk(tmp_path):
    """a 0.6 m box sampled from the library's own random initial points (~44k): identical to the restatement run from the same
    points, a plausible areal density, and no two samples of one face closer than particleRadius."""
    from oracle.oracle import BoundryOracle
    from wcsph_b200 import boundry as bd
    lo, hi = np.array([-0.3, 0.0, -0.3]), np.array([0.3, 0.6, 0.3])
    v = [(lo[0], lo[1], lo[2]), (hi[0], lo[1], lo[2]), (hi[0], hi[1], lo[2]), (lo[0], hi[1], lo[2]),
         (lo[0], lo[1], hi[2]), (hi[0], lo[1], hi[2]), (hi[0], hi[1], hi[2]), (lo[0], hi[1], hi[2])]
    f = [(1, 3, 2), (1, 4, 3), (5, 6, 7), (5, 7, 8), (1, 2, 6), (1, 6, 5), (4, 7, 3), (4, 8, 7), (1, 5, 8), (1, 8, 4), (2, 3, 7), (2, 7, 6)]
    os.chdir(tmp_path)
    _write_obj("tank.obj", v, f)
    n = bd.loadObj("tank.obj")
    assert n > 40000
    bd.init_point_set(seed=7)
    p0 = bd.fetch("pos")
    init_pos, init_id = p0[:n, :3].copy(), p0[:n, 3].view(np.int32).copy()
    assert init_id.min() >= 0 and init_id.max() < 12 and np.all(init_pos >= lo - 1e-5) and np.all(init_pos <= hi + 1e-5)
    assert len(np.unique(init_id)) == 12                               # every face drew points (area-weighted rejection)
    bd.gpu_bitonic_sort()
    bd.build_hmap()
    samples = bd.sample_all()
    o = BoundryOracle(bd._s["tri_normal"], init_pos, init_id, bd.min_point, bd.particleRadius)
    ref = o.run()
    assert len(samples) == len(ref) and np.array_equal(samples, ref)
    area = 6 * 0.36
    density = len(samples) / area                                       # the shipped clouds have ~1058 points / m^2
    assert 700 < density < 1500, density
    sel = bd.fetch("selected")[:len(samples)]
    ids = bd.fetch("pos")[:, 3].view(np.int32)[sel]
    for face in range(12):
        q = samples[ids == face].astype(np.float64)
        if len(q) > 1:
            d = np.sqrt(((q[:, None, :] - q[None, :, :]) ** 2).sum(-1)) + np.eye(len(q))
            assert d.min() >= bd.particleRadius * (1 - 1e-6)
    path = bd.export_obj("tank_boundry.obj", samples)
    assert sum(1 for line in open(path) if line.startswith("v ")) == len(samples)
