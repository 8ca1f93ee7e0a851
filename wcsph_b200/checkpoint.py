"""Checkpoint / restart of a solver's persistent state (the reference has none: its only
outputs are PNG frames and OBJ dumps).  State = the fields that survive a step, in the reference's insertion
order, + the time step and iteration counters; everything else is recomputed by the next step.

Works on z-slab ranks too: `save_state` assembles the global arrays (every rank contributes its rows; rank 0
writes), `load_state` re-partitions the liquids by the restored positions before it sets the other fields."""
import os
import tempfile

import numpy as np

_PERSISTENT = {
    "dfsph": ("pos", "vel", "omega", "vel_guess", "kappa", "kappa_v", "pressure"),
    "iisph": ("pos", "vel", "vel_guess", "pressure"),
    "pcisph": ("pos", "vel"),
    "sesph": ("pos", "vel"),
}


def _path(path):
    path = os.fspath(path)
    return path if path.endswith(".npz") else path + ".npz"      # np.savez would append it behind our back


def _gather(pd, name):
    """field in reference order; on z-slab ranks Field.to_numpy() assembles it (every rank fills its rows, summed over the ranks)."""
    return getattr(pd, name).to_numpy()


def save_state(module, path):
    """module: an initialised wcsph_b200 solver module (dfsph / iisph / pcisph / sesph).  Atomic: the file appears
    under its final name only when it is complete."""
    pd = module.particle_data
    out = {n: _gather(pd, n) for n in _PERSISTENT[pd.solver]}
    out["deltaT"] = pd.deltaT.to_numpy()
    out["iters"] = np.array([getattr(module, "vs_iter", 0), getattr(module, "dv_iter", 0), getattr(module, "pr_iter", 0)], dtype=np.int32)
    out["current_time"] = np.array([getattr(module, "current_time", 0.0)], dtype=np.float64)
    out["counts"] = np.array([pd.count, pd.liquid_count], dtype=np.int64)
    out["solver"] = np.frombuffer(pd.solver.encode(), dtype=np.uint8)
    out["bbox"] = np.concatenate([pd.minboundarynp[0], pd.maxboundarynp[0]]).astype(np.float32)
    out["gridR"] = np.array([pd.hash_grid.gridR], dtype=np.float64)
    path = _path(path)
    if pd.world_size > 1 and pd.rank != 0:
        return path
    fd, tmp = tempfile.mkstemp(suffix=".npz.tmp", dir=os.path.dirname(os.path.abspath(path)))
    try:
        with os.fdopen(fd, "wb") as f:
            np.savez(f, **out)
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise
    return path


def load_state(module, path):
    """restore into a module initialised on the SAME scene: solver, particle counts, bounding box and hash cell are verified."""
    pd = module.particle_data
    z = np.load(_path(path))
    if tuple(z["counts"]) != (pd.count, pd.liquid_count):
        raise ValueError("checkpoint is for %s particles, scene has %s" % (tuple(z["counts"]), (pd.count, pd.liquid_count)))
    if "solver" in z.files:
        name = bytes(z["solver"]).decode()
        if name != pd.solver:
            raise ValueError("checkpoint of solver '%s' loaded into '%s'" % (name, pd.solver))
        bbox = np.concatenate([pd.minboundarynp[0], pd.maxboundarynp[0]]).astype(np.float32)
        if not np.array_equal(z["bbox"], bbox) or float(z["gridR"][0]) != float(pd.hash_grid.gridR):
            raise ValueError("checkpoint is for another scene: bounding box / hash cell differ (%s, %s) vs (%s, %s)"
                             % (z["bbox"], z["gridR"], bbox, pd.hash_grid.gridR))
    if pd.world_size > 1:
        pd.reupload(z["pos"])            # re-home: the restored positions decide the owner of every liquid particle
    for n in _PERSISTENT[pd.solver]:
        if n == "pos" and pd.world_size > 1:
            continue
        getattr(pd, n).from_numpy(z[n])
    pd.deltaT.from_numpy(z["deltaT"])
    for k, n in enumerate(("vs_iter", "dv_iter", "pr_iter")):
        if hasattr(module, n):
            setattr(module, n, int(z["iters"][k]))
    if hasattr(module, "current_time"):
        module.current_time = float(z["current_time"][0])
    from . import _lib
    _lib.check(_lib.load().wcsph_set_iters(pd._ctx, int(z["iters"][0]), int(z["iters"][1]), int(z["iters"][2])))
