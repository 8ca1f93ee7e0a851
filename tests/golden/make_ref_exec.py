"""Reference-EXECUTED goldens: run the unmodified /root/reference solver scripts under the serial
Taichi shim (oracle/tishim) on small scenes and record what every kernel wrote.

    python tests/golden/make_ref_exec.py [sesph pcisph iisph dfsph] [--steps K] [--dim D]

writes tests/golden/ref_exec_<solver>.npz.  Runs only in the build container (it reads
/root/reference); the .npz files are committed and travel.  What is changed relative to running
`ti <solver>.py` -- all of it here, none of it in the reference's source text:

* scene size: the module constants `particleDimX/Y/Z` (and the canvas size) are substituted so that
  pure Python finishes in minutes; sesph / pcisph also get `boundary` shrunk so their generated
  lattice shell stays close to the liquid block.  dfsph / iisph read "model/box_boundry.obj"
  relative to the cwd: the cwd is a scratch directory that holds a SMALL open box in that place
  (`small_box_obj` below, committed inside the .npz as `solid_pos`).
* the GUI loop ends after K frames.
* pcisph `compute_nonpressure_force` is launched twice per step: its single parallel loop resets
  and accumulates rho[i] while reading rho[j] (Q24 -- a data race in Taichi, NaN on step 0 in any
  serial order).  The second launch reads a complete rho for every j, which is exactly the
  two-phase definition D-PCI that the oracle and the CUDA path implement.
* out-of-bounds accesses (Q7, Q12, Q15) read 0 / are dropped and are COUNTED in the .npz (`oob`).

Event stream: after every kernel (and before it, to catch host-side writes such as
`deltaT.from_numpy`, dfsph.py:129) every tracked field is compared with its previous snapshot;
changed fields are stored as `e<idx>_<field>`.  `events` lists (idx, kernel, [fields]).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "tishim"))


def small_box_obj(path, lo, hi, spacing):
    """open-top box of lattice points around [lo, hi] (floor + 4 walls), `v x y z` lines."""
    pts = []
    nx = int(round((hi[0] - lo[0]) / spacing)) + 1
    ny = int(round((hi[1] - lo[1]) / spacing)) + 1
    nz = int(round((hi[2] - lo[2]) / spacing)) + 1
    for a in range(nx):
        for b in range(ny):
            for cc in range(nz):
                if a in (0, nx - 1) or cc in (0, nz - 1) or b == 0:
                    pts.append((lo[0] + a * spacing, lo[1] + b * spacing, lo[2] + cc * spacing))
    with open(path, "w") as f:
        f.write("# small open box for the reference-executed goldens\n")
        for p in pts:
            f.write("v %.6f %.6f %.6f\n" % p)
    return np.array(pts, np.float64)


def scene_consts(solver, dim, img, dimz=None):
    c = {"particleDimX": dim, "particleDimY": dim, "particleDimZ": dimz or dim}
    if solver in ("sesph", "pcisph"):
        # shell of int(boundary/gridR)^3 lattice nodes on [-boundary/2, boundary/2]; liquid block sits at
        # x,z in [0, (dim-1)*0.05], y from -0.9.  boundary 2.0 keeps y=-1 as the floor 0.1 below the block.
        c["imgSize"] = img
    else:
        c["imgSizeX"] = img
        c["imgSizeY"] = img
    return c


class Recorder:
    def __init__(self, solver, kick=0.0, surface=False):
        self.solver = solver
        self.kick = kick
        self.surface = surface
        self.fields = {}      # label -> Field
        self.prev = {}
        self.events = []
        self.arrays = {}
        self.steps = []
        self.mod = None
        self.busy = False
        self.t0 = time.time()

    def discover(self):
        import taichi as ti
        mod = sys.modules[self.solver]
        self.mod = mod
        pd = getattr(mod, "particle_data", None)
        if pd is None:
            return False
        found = {}
        for k, v in vars(mod).items():
            if isinstance(v, ti.Field):
                found[k] = v
        for k, v in vars(pd).items():
            if isinstance(v, ti.Field) and v.arr is not None:
                found.setdefault(k, v)
        hg = pd.hash_grid
        for k in ("gridCount", "grid", "neighborCount", "neighbor"):
            found["hg_" + k] = getattr(hg, k)
        if self.surface:
            mc = pd.mc_grid
            for k in ("gridCount", "grid", "surface_value", "vertex_count", "triangle"):
                found["mc_" + k] = getattr(mc, k)
        cv = getattr(mod, "sph_canvas", None)
        if cv is not None:
            found["canvas_img"] = cv.img
            found["canvas_depth"] = cv.depth
            found["canvas_view"] = cv.view
            found["canvas_proj"] = cv.proj
        self.fields = found
        return True

    def snapshot(self, kernel):
        if not self.fields and not self.discover():
            return
        changed = []
        idx = len(self.events)
        for label, f in self.fields.items():
            if label in ("hg_grid", "hg_neighbor", "hg_gridCount", "hg_neighborCount") and kernel != "update_grid":
                continue
            if label.startswith("mc_") and kernel not in ("update_grid", "cal_surface_point", "cal_surface_point_anistropic", "marching_cube"):
                continue
            if label in ("color", "color_grad", "pos_avr", "G") and kernel not in ("compute_color_map", "cal_anistropic_kernel"):
                continue
            if label.startswith("canvas_") and kernel not in ("draw_particle", "<host>"):
                continue
            if label in ("canvas_img", "canvas_depth") and kernel != "draw_particle":
                continue
            a = f.arr
            p = self.prev.get(label)
            if p is None or p.shape != a.shape or not np.array_equal(p, a, equal_nan=True):
                self.prev[label] = a.copy()
                changed.append(label)
                self._store(idx, label, a)
        if changed or kernel != "<host>":
            self.events.append((idx, kernel, changed))

    def _store(self, idx, label, a):
        key = "e%d_%s" % (idx, label)
        if label == "hg_neighbor":
            cnt = self.fields["hg_neighborCount"].arr
            cap = a.shape[1]
            rows = [a[i, :min(int(cnt[i]), cap)] for i in range(a.shape[0])]
            self.arrays[key] = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
        elif label == "mc_triangle":
            n = int(self.fields["mc_vertex_count"].arr[0])
            self.arrays[key] = a[:min(n, a.shape[0])].copy()
        elif label == "mc_grid":
            self.arrays[key] = a.copy()
        elif label == "hg_grid":
            cnt = self.fields["hg_gridCount"].arr
            rows = [a[i, :int(cnt[i])] for i in range(a.shape[0])]
            self.arrays[key] = np.concatenate(rows).astype(np.int32)
        else:
            self.arrays[key] = a.copy()

    def hook(self, name, owner, phase):
        import taichi as ti
        if self.busy:
            return
        if phase == "pre":
            self.snapshot("<host>")
            return
        if name == "reset_param" and self.kick and (self.fields or self.discover()):
            # developed-flow variant: a deterministic velocity field injected into the STATE right after the
            # reference's own reset_param (no source change): particles cross cell faces within the recorded steps
            vel = self.fields["vel"]
            p = np.array(self.mod.particle_data.point_list, dtype=np.float64)[:vel.arr.shape[0]]
            c0 = p.mean(axis=0)
            q = (p - c0) * 9.0
            v = np.stack([np.sin(q[:, 1]) + 0.5 * np.cos(q[:, 2]), -0.6 + 0.4 * np.sin(q[:, 0] + q[:, 2]),
                          np.cos(q[:, 0]) - 0.5 * np.sin(q[:, 1])], axis=1) * self.kick
            vel.from_numpy(v.astype(np.float32))
        if self.solver == "pcisph" and name == "compute_nonpressure_force":
            self.busy = True          # D-PCI: second launch reads complete rho (see module docstring)
            try:
                self.mod.compute_nonpressure_force()
            finally:
                self.busy = False
        self.snapshot(name)

    def on_show(self, gui):
        m = self.mod
        self.snapshot("<host>")
        it = {k: int(getattr(m, k)) for k in ("vs_iter", "dv_iter", "pr_iter") if hasattr(m, k)}
        it["deltaT"] = float(m.deltaT.arr[0])
        it["event_end"] = len(self.events)
        self.steps.append(it)
        print("[%s] step %d done: %s  (%.0f s)" % (self.solver, len(self.steps), it, time.time() - self.t0), flush=True)


def run(solver, steps, dim, img, out, kick=0.0, dimz=None, surface=False):
    import taichi as ti
    from loader import RefLoader

    consts = {solver: scene_consts(solver, dim, img, dimz)}
    if solver in ("sesph", "pcisph"):
        consts[solver]["boundary"] = 1.9      # floor of the lattice shell at y = -0.95, one particle spacing below the block
    work = tempfile.mkdtemp(prefix="refexec_")
    os.makedirs(os.path.join(work, "model"))
    os.makedirs(os.path.join(work, "out"))
    os.symlink(os.path.join(REF, "MCData.txt"), os.path.join(work, "MCData.txt"))
    d = 0.05
    if solver == "dfsph":
        half = dim / 2 * d
        lo = (-half - 0.5 * d, 0.2 - d, -half - 0.5 * d)
        hi = (half + 0.5 * d, 0.2 + dim * d, (dimz or dim) * d - half + 0.5 * d)
    else:   # iisph block: x,z from -0.025, y from 0.1
        lo = (-0.025 - d, 0.1 - d, -0.025 - d)
        hi = (-0.025 + dim * d, 0.1 + dim * d, -0.025 + (dimz or dim) * d)
    solid = small_box_obj(os.path.join(work, "model", "box_boundry.obj"), lo, hi, 0.03)

    rec = Recorder(solver, kick, surface)
    ti.trace_hook[0] = rec.hook
    ti.GUI.max_frames = steps
    ti.GUI.on_show = rec.on_show
    ti.oob_log.clear()
    loader = RefLoader(REF, consts).install()
    cwd = os.getcwd()
    os.chdir(work)
    try:
        try:
            __import__(solver)
        except SystemExit:
            pass
        if surface:
            # SURVEY 8(f) N2, both branches of MarchingCubeGrid.export_surface (:137-157), called on the state the K steps left:
            # the active one (update_grid -> cal_surface_point -> marching_cube) and the one the reference has commented out
            # (compute_color_map, cal_anistropic_kernel -> cal_surface_point_anistropic -> marching_cube)
            pd = rec.mod.particle_data
            mc = pd.mc_grid
            rec.fields = {}
            rec.discover()
            mc.update_grid()
            mc.cal_surface_point()
            mc.marching_cube()
            pd.compute_color_map()
            pd.cal_anistropic_kernel()
            mc.update_grid()
            mc.cal_surface_point_anistropic()
            mc.marching_cube()
    finally:
        os.chdir(cwd)
        loader.remove()
        ti.trace_hook[0] = None
        ti.GUI.on_show = None
    mod = rec.mod
    pd = mod.particle_data
    init_pos = np.array(pd.point_list, dtype=np.float32)
    meta = {
        "solver": solver, "kick": kick, "steps": steps, "dim": dim, "dimz": dimz or dim, "img": img, "consts": consts[solver],
        "events": [[i, k, f] for i, k, f in rec.events],
        "step_info": rec.steps,
        "oob": [[str(k[0]), str(k[1]), k[2], v] for k, v in sorted(ti.oob_log.items(), key=str)],
        "rewritten": {n: r for n, r in loader.loaded},
        "count": int(pd.count), "liquid_count": int(pd.liquid_count),
        "blockSize": [int(v) for v in pd.hash_grid.blockSize.arr[0]],
        "min_boundary": [float(v) for v in pd.hash_grid.min_boundary.arr[0]],
        "max_boundary": [float(v) for v in pd.hash_grid.max_boundary.arr[0]],
        "hash_gridR": float(pd.hash_grid.gridR),
        "mc": {"gridR": float(pd.mc_grid.gridR), "searchR": float(pd.mc_grid.searchR), "maxInGrid": int(pd.mc_grid.maxInGrid),
               "block": [int(v) for v in pd.mc_grid.blocknp[0]], "min_boundary": [float(v) for v in pd.mc_grid.minboundarynp[0]],
               "isolevel": float(pd.mc_grid.isolevel), "liqiudMass": float(pd.liqiudMass)},
        "pci_coff": float(getattr(mod, "pci_coff", 0.0)),
        "reference_commit": "37f79c2",
    }
    arrays = dict(rec.arrays)
    arrays["init_pos"] = init_pos
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out, **arrays)
    print("[%s] wrote %s: %d events, %d arrays, %.1f MB, oob=%s" % (
        solver, out, len(rec.events), len(arrays), os.path.getsize(out) / 1e6, meta["oob"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("solvers", nargs="*", default=["sesph", "pcisph", "iisph", "dfsph"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dim", type=int, default=8)
    ap.add_argument("--img", type=int, default=64)
    ap.add_argument("--suffix", default="")
    ap.add_argument("--dimz", type=int, default=None)
    ap.add_argument("--surface", action="store_true", help="after the steps: both surface-reconstruction branches (SURVEY 8f N2)")
    ap.add_argument("--kick", type=float, default=0.0, help="amplitude (m/s) of the velocity field injected after reset_param")
    a = ap.parse_args()
    for s in a.solvers:
        run(s, a.steps, a.dim, a.img, os.path.join(HERE, "ref_exec_%s%s.npz" % (s, a.suffix)), a.kick, a.dimz, a.surface)


if __name__ == "__main__":
    main()
