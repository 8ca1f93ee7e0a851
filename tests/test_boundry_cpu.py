"""SURVEY 8(f) N3 -- boundry.py (parallel Poisson-disk boundary sampler).  The restatement oracle/boundry_oracle.c against
tests/golden/ref_exec_boundry.npz, which the UNMODIFIED /root/reference/boundry.py produced under the serial Taichi shim
(tests/golden/make_ref_exec_boundry.py): every stage bit-exact from the initial point set the reference drew."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_exec_boundry.npz")


def load_golden():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["meta"]).decode())


def test_boundry_oracle_reproduces_reference_executed_stages():
    from oracle.oracle import BoundryOracle
    z, meta = load_golden()
    n = meta["numInitialPoints"]
    assert meta["reference_commit"] == "37f79c2" and n > 4000
    o = BoundryOracle(z["tri_normal"], z["init_pos"][:n], z["init_id"][:n], meta["min_point"], meta["particleRadius"])
    assert (o.p.padding, o.p.hash_size, o.p.phase_vec_max) == (meta["padding_num"], meta["hash_map_size"], meta["phase_vec_max"])
    o.cells()
    assert np.array_equal(o.cell, z["init_cell"])                                   # boundry.py:245-247
    o.sort()                                                                        # :208-219: the permutation of the bitonic network
    assert np.array_equal(o.cell, z["sorted_cell"]) and np.array_equal(o.pos, z["sorted_pos"]) and np.array_equal(o.id, z["sorted_id"])
    assert o.build_hmap() == int(z["hash_count"][0])                                # :250-271
    assert np.array_equal(o.start_index, z["hmap_start_index"]) and np.array_equal(o.hcell, z["hmap_cell"])
    assert np.array_equal(o.hash_trace, z["hash_trace"])
    assert np.array_equal(o.phase_group_count, z["phase_group_count"]) and np.array_equal(o.phase_group, z["phase_group"])
    order = BoundryOracle.launch_order()
    assert len(order) == len(z["launch"]) == 269 and order[0] == (1, 0) and (0, 0) not in order and order[-1] == (26, 9)
    counts = [o.sample_launch(pg, t) for pg, t in order]                            # :390-407, launch by launch
    assert np.array_equal(np.array(counts), z["launch"])
    assert o.n_sample == meta["sample_count"] == meta["obj_lines"] and o.n_sample > 200
    assert np.array_equal(o.possion_sample[:o.n_sample], z["possion_sample"])
    assert np.array_equal(o.sample_count, z["hmap_sample_count"])
    used = o.sample_count > 0
    assert all(np.array_equal(o.sample[h, :o.sample_count[h]], z["hmap_sample"][h, :o.sample_count[h]]) for h in np.nonzero(used)[0])


def test_boundry_module_surface_and_launch_order():
    """the product module keeps the reference's names and its (phase, trial) launch sequence"""
    from oracle.oracle import BoundryOracle
    from wcsph_b200 import boundry
    for name in ("particleRadius", "gridR", "phase_block_size", "hash_sample_size", "get_pot_num", "loadObj", "init_point_set",
                 "gpu_bitonic_sort", "build_hmap", "detect_hmap", "possion_disk_sample"):
        assert hasattr(boundry, name)
    assert boundry.launch_order() == BoundryOracle.launch_order()
    assert boundry.get_pot_num(5003) == 4096 and boundry.get_pot_num(4096) == 2048 and boundry.gridR == 0.025 / 3 ** 0.5
