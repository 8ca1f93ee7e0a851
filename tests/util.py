"""Shared helpers of the parity tests: same scene into the oracle (CPU restatement) and the
CUDA engine (through the reference-shaped Python surface -> C ABI)."""
import importlib

import numpy as np

from wcsph_b200 import scenes

TOL = 1e-4      # north-star tolerance: per-step fields within 1e-4 (scale-normalised)


def scene(solver, kind="asshipped", dims=None):
    if kind == "asshipped":
        return getattr(scenes, "scene_" + solver)()
    if kind == "dam":
        nx, ny, nz = dims or (16, 16, 16)
        return scenes.dam_break(nx, ny, nz, jitter=True, config_id=2)
    if kind == "dam32":          # the C2 generator at 32^3 for alias coverage (32,768 liquid + 12,696 boundary)
        return scenes.dam_break(32, 32, 32, jitter=True, config_id=2)
    if kind == "dam_lattice":
        nx, ny, nz = dims or (16, 16, 16)
        return scenes.dam_break(nx, ny, nz, jitter=False)
    raise ValueError(kind)


def make_oracle(solver, pts, nl, **over):
    from oracle.oracle import Oracle
    return Oracle(solver, pts, nl, threads=8, **over)


def make_engine(solver, pts, nl, **kw):
    m = importlib.import_module("wcsph_b200." + solver)
    m = importlib.reload(m)            # fresh module globals per scene
    m.init_scene(pts, nl, **kw)
    m.reset_param()
    return m


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(max|b|, floor): scale-normalised infinity norm."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, floor, 1e-30)
    return float(np.max(np.abs(a - b))) / scale if b.size else 0.0


def assert_close(name, a, b, tol=TOL, floor=0.0):
    assert np.all(np.isfinite(a)), "%s: non-finite values from the CUDA path" % name
    e = rel_err(a, b, floor)
    assert e <= tol, "%s: scale-normalised error %.3e > %.1e" % (name, e, tol)
    return e


def eng_field(m, name):
    return getattr(m.particle_data, name).to_numpy()


def eng_scalar(m, name):
    return float(getattr(m.particle_data, name).to_numpy()[0])
