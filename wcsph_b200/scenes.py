"""Scene generators: the as-shipped particle layouts of the four reference scripts
and the synthetic dam-break of SURVEY.md 8(d).  Host-side numpy only (this is
input generation, not the hot path).

Every generator returns (points float64 (N,3) in insertion order, liquid_count).
Liquid particles come first, solids after -- the ordering contract every
reference kernel's `j < particleLiquidNum` relies on (dfsph.py:258).
"""
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_obj_vertices(filename):
    """ParticleData.py:130-138 -- OBJ 'v' lines only, '#' comments skipped."""
    pts = []
    with open(filename, "r") as f:
        for line in f:
            if line.startswith("#"):
                continue
            values = line.split()
            if not values:
                continue
            if values[0] == "v":
                pts.append(list(map(float, values[1:4])))
    return np.asarray(pts, dtype=np.float64).reshape(-1, 3)


def load_boundary(name_or_path):
    """An OBJ path, or the name of a committed fixture (tests/golden/<name>.npy)."""
    if os.path.exists(name_or_path):
        if name_or_path.endswith(".npy"):
            return np.load(name_or_path).astype(np.float64)
        return load_obj_vertices(name_or_path)
    base = os.path.splitext(os.path.basename(name_or_path))[0]
    cand = os.path.join(_GOLDEN, base + ".npy")
    if os.path.exists(cand):
        return np.load(cand).astype(np.float64)
    raise FileNotFoundError(name_or_path)


def dfsph_liquid_block(particleRadius=0.025, dims=(20, 20, 20)):
    """dfsph.py:66-73."""
    dx, dy, dz = dims
    n = dx * dy * dz
    ZxY = dz * dy
    dis = particleRadius * 2.0
    i = np.arange(n)
    x = (i // ZxY - dx / 2).astype(np.float64) * dis + dis * 0.5
    y = ((i % ZxY) // dz).astype(np.float64) * dis + 0.2
    z = (i % dz - dz / 2).astype(np.float64) * dis + dis * 0.5
    return np.stack([x, y, z], axis=1)


def scene_dfsph(boundary="box_boundry", particleRadius=0.025, dims=(20, 20, 20)):
    """dfsph.py:59-82 with init_particle("model/box_boundry.obj") (dfsph.py:597)."""
    liq = dfsph_liquid_block(particleRadius, dims)
    sol = load_boundary(boundary)
    return np.concatenate([liq, sol], axis=0), liq.shape[0]


def _lattice_shell(boundary=2.0, gridR=0.05):
    """sesph.py:66-90 / pcisph.py:117-141 solid shell."""
    invGridR = 1.0 / gridR
    blockSize = int(boundary * invGridR)
    A = boundary / (float(blockSize) - 1.0)
    B = -0.5 * boundary
    g = np.arange(blockSize)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")     # i = ix*bs*bs + iy*bs + iz
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    m = (ix == 0) | (iy == 0) | (iz == 0) | (ix == blockSize - 1) | (iy == blockSize - 1) | (iz == blockSize - 1)
    return np.stack([A * ix[m].astype(np.float64) + B, A * iy[m].astype(np.float64) + B,
                     A * iz[m].astype(np.float64) + B], axis=1)


def scene_sesph(particleRadius=0.025, dims=(20, 20, 20)):
    """sesph.py:66-92 (identical generator in pcisph.py:117-143)."""
    gridR = particleRadius * 2.0
    dx, dy, dz = dims
    n = dx * dy * dz
    ZxY = dz * dy
    i = np.arange(n)
    liq = np.stack([(i // ZxY).astype(np.float64) * gridR,
                    ((i % ZxY) // dz).astype(np.float64) * gridR - 0.9,
                    (i % dz).astype(np.float64) * gridR], axis=1)
    return np.concatenate([liq, _lattice_shell(2.0, gridR)], axis=0), n


scene_pcisph = scene_sesph


def scene_iisph(boundary="box_boundry", particleRadius=0.025, dims=(20, 20, 20)):
    """iisph.py:99-112 with init_particle("model/box_boundry.obj") (iisph.py:411)."""
    dx, dy, dz = dims
    n = dx * dy * dz
    ZxY = dz * dy
    dis = particleRadius * 2.0
    i = np.arange(n)
    liq = np.stack([(i // ZxY).astype(np.float64) * dis - particleRadius,
                    ((i % ZxY) // dz).astype(np.float64) * dis + 0.1,
                    (i % dz).astype(np.float64) * dis - particleRadius], axis=1)
    sol = load_boundary(boundary)
    return np.concatenate([liq, sol], axis=0), n


def dam_break(nx, ny, nz, particleRadius=0.025, jitter=False, config_id=0):
    """SURVEY.md 8(d) synthetic scene.  Liquid index i = ix*ny*nz + iy*nz + iz at
    ((ix+1)d, (iy+1)d, (iz+1)d); one-layer lattice shell on the box of nodes
    bx=2nx+2, by=floor(1.5ny)+2, bz=nz+2 at spacing d from the origin, appended after
    the liquid.  Optional jitter: default_rng(1234+config_id), uniform +-0.05d."""
    d = 2.0 * particleRadius
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    liq = np.stack([(ix.ravel() + 1) * d, (iy.ravel() + 1) * d, (iz.ravel() + 1) * d], axis=1).astype(np.float64)
    if jitter:
        rng = np.random.default_rng(1234 + config_id)
        liq = liq + rng.uniform(-0.05 * d, 0.05 * d, size=liq.shape)
    bx, by, bz = 2 * nx + 2, int(1.5 * ny) + 2, nz + 2
    faces = []
    gx, gy, gz = np.arange(bx), np.arange(by), np.arange(bz)
    # shell = nodes with any coordinate on the box surface, enumerated x-major like sesph.py:84-90
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij", sparse=True)
    m = (X == 0) | (Y == 0) | (Z == 0) | (X == bx - 1) | (Y == by - 1) | (Z == bz - 1)
    sx, sy, sz = np.nonzero(m)
    del faces
    sol = np.stack([sx * d, sy * d, sz * d], axis=1).astype(np.float64)
    return np.concatenate([liq, sol], axis=0), liq.shape[0]
