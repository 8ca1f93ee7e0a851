"""Generates the committed input fixtures from the reference's own data files.

Run in the build container (where /root/reference exists):
    python tests/golden/make_fixtures.py
Outputs (float32, the precision ParticleData.setup_data_cpu uploads, ParticleData.py:182):
    box_boundry.npy  <- model/box_boundry.obj  (dfsph.py:597, iisph.py:411 boundary cloud)
    liqiud.npy       <- model/liqiud.obj       (dump of dfsph.py:70-73, ParticleData.py:101-108)
    anchors.json     <- numbers the reference lets us evaluate without Taichi
    mc_tables.npz    <- MCData.txt, the marching-cubes case tables MarchingCubeGrid.setup_grid_cpu loads
                        (MarchingCubeGrid.py:80-94): edgetable i32[256], tritable i32[256,16]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"


def obj_vertices(path):
    pts = []
    for line in open(path):
        v = line.split()
        if v and v[0] == "v":
            pts.append([float(x) for x in v[1:4]])
    return np.asarray(pts, dtype=np.float64)


def pci_coff_from_reference_source():
    """Execute pcisph.py:74-115 verbatim (pure numpy) by extracting the two defs with ast."""
    import ast
    src = open(os.path.join(REF, "pcisph.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("CpuGradW", "GetPciCoff")]
    consts = {}
    for n in tree.body:      # module-level float constants they use (pcisph.py:24-45)
        if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name):
            name = n.targets[0].id
            if name in ("particleRadius", "gridR", "searchR", "pi", "h3", "m_k", "m_l"):
                consts[name] = eval(compile(ast.Expression(n.value), "pcisph", "eval"), {}, consts)
    env = dict(consts)
    env["np"] = np
    exec(compile(ast.Module(keep, []), "pcisph", "exec"), env)
    return float(env["GetPciCoff"]())


def mc_tables():
    """MarchingCubeGrid.py:80-94 applied to the reference's MCData.txt."""
    edge, tri = [], []
    for li, line in enumerate(open(os.path.join(REF, "MCData.txt"))):
        vals = [v.strip() for v in line.strip().split(",") if v.strip()]
        if li < 32:
            edge += [int(v, 16) for v in vals]
        elif vals:
            tri.append([int(v) for v in vals])
    return np.asarray(edge, np.int32), np.asarray(tri, np.int32)


def canvas_mc_goldens():
    """Canvas / marching-cubes goldens from the CPU restatement on the as-shipped DFSPH scene after 3 oracle steps
    (static_cam(0,1,0), dfsph-style draw_particle; MCGrid(particleR, 4, 512)).  Like oracle_steps.npz they pin the
    ORACLE against regressions; they are only as authoritative as the restatement (parity unpinned)."""
    import hashlib
    from oracle import oracle
    from wcsph_b200 import scenes
    from wcsph_b200.Canvas import Canvas
    pts, nl = scenes.scene_dfsph()
    o = oracle.Oracle("dfsph", pts, nl, threads=1)
    for _ in range(3):
        o.step()
    cv = Canvas(512, 512)
    cv.static_cam(0.0, 1.0, 0.0)
    img, depth = oracle.canvas_draw_particle(o.field("pos"), nl, cv.view[0], cv.proj[0], 512, 512, 1)
    mc = oracle.McOracle(pts, nl, threads=1)
    mc.update_grid(o.field("pos"))
    sv = mc.cal_surface_point(o.field("rho")).copy()
    e, t = mc_tables()
    n, v = mc.marching_cube(e, t)
    return {
        "canvas_lit_pixels": int((img[:, :, 0] > 0).sum()), "canvas_white_pixels": int((img[:, :, 0] == 1.0).sum()),
        "canvas_img_sha256": hashlib.sha256(img.tobytes()).hexdigest(), "canvas_depth_sha256": hashlib.sha256(depth.tobytes()).hexdigest(),
        "mc_block": [int(x) for x in mc.block], "mc_nodes_above_iso": int((sv > 0.5).sum()),
        "mc_surface_sha256": hashlib.sha256(sv.tobytes()).hexdigest(),
        "mc_vertex_count": int(n), "mc_mesh_sha256": hashlib.sha256(v.tobytes()).hexdigest(),
    }


def oracle_goldens():
    """Per-step goldens from the CPU restatement (oracle/), single thread, for the as-shipped scenes.
    They pin the ORACLE against regressions and give the GPU tests committed vectors to hit; they are only
    as authoritative as the restatement (parity unpinned: no Taichi here, see oracle/wcsph_oracle.h)."""
    from oracle.oracle import Oracle
    from wcsph_b200 import scenes
    out = {}
    for solver, steps in (("dfsph", 3), ("sesph", 3), ("iisph", 2), ("pcisph", 2)):
        pts, nl = getattr(scenes, "scene_" + solver)()
        o = Oracle(solver, pts, nl, threads=1)
        for _ in range(steps):
            o.step()
        out[solver + "_rho"] = o.field("rho").astype(np.float32).copy()
        out[solver + "_pos"] = o.field("pos")[:nl].astype(np.float32).copy()
        out[solver + "_vel"] = o.field("vel").astype(np.float32).copy()
        out[solver + "_iters"] = np.array([o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter"), steps], dtype=np.int32)
        out[solver + "_dt"] = np.array([o.get("deltaT")], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "oracle_steps.npz"), **out)
    return {k: v.shape for k, v in out.items()}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "oracle":
        print(oracle_goldens())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "canvas_mc":
        g = canvas_mc_goldens()
        json.dump(g, open(os.path.join(HERE, "canvas_mc_goldens.json"), "w"), indent=1)
        print(g)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mc":
        e, t = mc_tables()
        np.savez_compressed(os.path.join(HERE, "mc_tables.npz"), edgetable=e, tritable=t)
        print("mc_tables.npz", e.shape, t.shape)
        sys.exit(0)
    box = obj_vertices(os.path.join(REF, "model", "box_boundry.obj"))
    liq = obj_vertices(os.path.join(REF, "model", "liqiud.obj"))
    np.save(os.path.join(HERE, "box_boundry.npy"), box.astype(np.float32))
    np.save(os.path.join(HERE, "liqiud.npy"), liq.astype(np.float32))
    anchors = {
        "pci_coff": pci_coff_from_reference_source(),
        "box_boundry_count": int(box.shape[0]),
        "liqiud_count": int(liq.shape[0]),
        "box_boundry_bbox": [box.min(0).tolist(), box.max(0).tolist()],
    }
    json.dump(anchors, open(os.path.join(HERE, "anchors.json"), "w"), indent=1)
    print(anchors)
