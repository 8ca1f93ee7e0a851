"""Reference-EXECUTED golden of the boundary pre-processing tool: runs the unmodified /root/reference/boundry.py
(parallel Poisson-disk sampling of a triangle mesh: random initial points, bitonic sort by cell, hash map, 27 phase groups,
10 trials) under the serial Taichi shim on a small closed box mesh and records every stage.

    python tests/golden/make_ref_exec_boundry.py        -> tests/golden/ref_exec_boundry.npz

Only in the build container (reads /root/reference).  The script's `ti.random()` draws come from the shim's seeded generator; the
golden stores the initial point set it produced, so that the oracle and the CUDA path can start from the SAME points -- from there
on the algorithm is deterministic (serial order defines the two places where the reference's parallel loops race: which of two
colliding cells keeps a hash slot, and the order of appends).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "tishim"))


def box_obj(path, lo, hi):
    """closed box, 8 vertices / 12 triangles, outward normals"""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)]
    f = [(1, 3, 2), (1, 4, 3), (5, 6, 7), (5, 7, 8), (1, 2, 6), (1, 6, 5), (4, 7, 3), (4, 8, 7), (1, 5, 8), (1, 8, 4), (2, 3, 7), (2, 7, 6)]
    with open(path, "w") as fo:
        for p in v:
            fo.write("v %.6f %.6f %.6f\n" % p)
        for t in f:
            fo.write("f %d %d %d\n" % t)
    return np.array(v, np.float64), np.array(f, np.int32)


if __name__ == "__main__":
    import taichi as ti
    from loader import RefLoader
    work = tempfile.mkdtemp(prefix="refexec_boundry_")
    verts, faces = box_obj(os.path.join(work, "box.obj"), (-0.11, 0.0, -0.08), (0.13, 0.17, 0.12))
    rec = {"launch": []}
    holder = {}

    def hook(name, owner, phase):
        if phase != "post":
            return
        m = sys.modules.get("boundry")
        if m is None:
            return
        holder["m"] = m
        if name == "init_point_set":
            rec["init_pos"] = m.init_pos.arr.copy()
            rec["init_id"] = m.init_id.arr.copy()
            rec["init_cell"] = m.init_cell.arr.copy()
        elif name == "gpu_merge":
            rec["sorted_pos"] = m.init_pos.arr.copy()
            rec["sorted_id"] = m.init_id.arr.copy()
            rec["sorted_cell"] = m.init_cell.arr.copy()
        elif name == "build_hmap":
            rec["hmap_start_index"] = m.hMap.start_index.arr.copy()
            rec["hmap_cell"] = m.hMap.cell.arr.copy()
            rec["hash_trace"] = m.hash_trace.arr.copy()
            rec["phase_group_count"] = m.phase_group_count.arr.copy()
            rec["phase_group"] = m.phase_group.arr.copy()
            rec["hash_count"] = m.hash_count_gpu.arr.copy()
        elif name == "possion_disk_sample":
            rec["launch"].append(int(m.possion_sample_count.arr[0]))

    ti.trace_hook[0] = hook
    ti.GUI.max_frames = 27 * 10 + 3
    ti.seed(20240611)
    ti.oob_log.clear()
    loader = RefLoader(REF, {"boundry": {"imgSize": 32}}).install()
    cwd = os.getcwd()
    os.chdir(work)
    try:
        __import__("boundry")
    except SystemExit:
        pass
    finally:
        os.chdir(cwd)
        ti.trace_hook[0] = None
    m = holder["m"]
    n = int(m.possion_sample_count.arr[0])
    out = dict(rec)
    out["launch"] = np.array(rec["launch"], np.int32)
    out["possion_sample"] = m.possion_sample.arr[:n].copy()
    out["hmap_sample_count"] = m.hMap.sample_count.arr.copy()
    out["hmap_sample"] = m.hMap.sample.arr.copy()
    out["tri_vertices"] = m.tri_vertices.arr.copy()
    out["tri_normal"] = m.tri_normal.arr.copy()
    out["tri_area"] = m.tri_area.arr.copy()
    obj_lines = open(os.path.join(work, "box_boundry.obj")).read().splitlines()
    meta = {"numInitialPoints": int(m.numInitialPoints), "padding_num": int(m.padding_num), "phase_vec_max": int(m.phase_vec_max),
            "hash_map_size": int(m.hash_map_size), "faceNum": int(m.faceNum), "particleRadius": float(m.particleRadius),
            "gridR": float(m.gridR), "min_point": [float(v) for v in m.min_point.e], "max_point": [float(v) for v in m.max_point.e],
            "totalArea": float(m.totalArea), "maxArea": float(m.maxArea), "sample_count": n, "obj_lines": len(obj_lines),
            "oob": [[str(k[0]), str(k[1]), k[2], v] for k, v in sorted(ti.oob_log.items(), key=str)],
            "mesh": {"vertices": verts.tolist(), "faces": faces.tolist()}, "reference_commit": "37f79c2",
            "launch_order": "possion_disk_sample(phase, trial): trial 0 phases 1..26, then trials 1..9 phases 0..26 (boundry.py:421-457)"}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    loader.remove()
    path = os.path.join(HERE, "ref_exec_boundry.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "initial points", meta["numInitialPoints"], "samples", n, "launches", len(rec["launch"]), "oob", meta["oob"])
