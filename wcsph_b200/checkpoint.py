"""Checkpoint / restart of a solver's persistent state (the reference has none: its only
outputs are PNG frames and OBJ dumps).  State = the fields that survive a step, in the reference's insertion
order, + the time step and iteration counters; everything else is recomputed by the next step."""
import numpy as np

_PERSISTENT = {
    "dfsph": ("pos", "vel", "omega", "vel_guess", "kappa", "kappa_v", "pressure"),
    "iisph": ("pos", "vel", "vel_guess", "pressure"),
    "pcisph": ("pos", "vel"),
    "sesph": ("pos", "vel"),
}


def save_state(module, path):
    """module: an initialised wcsph_b200 solver module (dfsph / iisph / pcisph / sesph)."""
    pd = module.particle_data
    out = {n: getattr(pd, n).to_numpy() for n in _PERSISTENT[pd.solver]}
    out["deltaT"] = pd.deltaT.to_numpy()
    out["iters"] = np.array([getattr(module, "vs_iter", 0), getattr(module, "dv_iter", 0), getattr(module, "pr_iter", 0)], dtype=np.int32)
    out["current_time"] = np.array([getattr(module, "current_time", 0.0)], dtype=np.float64)
    out["counts"] = np.array([pd.count, pd.liquid_count], dtype=np.int64)
    np.savez(path, **out)


def load_state(module, path):
    """restore into a module initialised on the SAME scene (same particle counts and boundary)."""
    pd = module.particle_data
    z = np.load(path)
    if tuple(z["counts"]) != (pd.count, pd.liquid_count):
        raise ValueError("checkpoint is for %s particles, scene has %s" % (tuple(z["counts"]), (pd.count, pd.liquid_count)))
    for n in _PERSISTENT[pd.solver]:
        getattr(pd, n).from_numpy(z[n])
    pd.deltaT.from_numpy(z["deltaT"])
    for k, n in enumerate(("vs_iter", "dv_iter", "pr_iter")):
        if hasattr(module, n):
            setattr(module, n, int(z["iters"][k]))
    if hasattr(module, "current_time"):
        module.current_time = float(z["current_time"][0])
    from . import _lib
    _lib.check(_lib.load().wcsph_set_iters(pd._ctx, int(z["iters"][0]), int(z["iters"][1]), int(z["iters"][2])))
