"""z-slab partition of the hash grid over the GPUs of one box.  Host-side numpy;
every rank evaluates it on the same scene and gets the same plan, so no communication is needed."""
import numpy as np


def cell_z(points, min_boundary, gridR):
    """cell layer of each point, with the arithmetic of HashGrid.py:68 (f32 subtract, f32 multiply by
    float(1/gridR), truncation toward zero)."""
    z = np.asarray(points, dtype=np.float64)[:, 2].astype(np.float32)
    inv = np.float32(1.0 / gridR)
    return ((z - np.float32(min_boundary[2])) * inv).astype(np.int32)


def block_size_z(min_boundary, max_boundary, gridR):
    """HashGrid.py:47: int((max - min) / gridR + 1) with an np.float32 difference."""
    return int(float(np.float32(max_boundary[2]) - np.float32(min_boundary[2])) / gridR + 1)


def z_slabs(liquid_points, min_boundary, max_boundary, gridR, world_size, min_layers=4):
    """Contiguous layer ranges [z_lo[r], z_hi[r]) with (nearly) equal liquid counts.

    Returns dict(z_lo, z_hi, counts, cap_own, cap_ghost).  cap_own / cap_ghost are the slot budgets a
    rank reserves: 1.5x the largest slab and 2x the largest two-layer boundary (+ slack), because the
    collapse moves particles between slabs."""
    bz = block_size_z(min_boundary, max_boundary, gridR)
    cz = np.clip(cell_z(liquid_points, min_boundary, gridR), 0, bz - 1)
    hist = np.bincount(cz, minlength=bz).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    if bz < world_size * min_layers:
        raise ValueError("grid has %d z layers: too thin for %d slabs of >= %d layers" % (bz, world_size, min_layers))
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        z = int(np.searchsorted(cum, target, side="left"))
        z = max(z, cuts[-1] + min_layers)
        z = min(z, bz - (world_size - r) * min_layers)
        cuts.append(z)
    cuts.append(bz)
    z_lo, z_hi = cuts[:-1], cuts[1:]
    counts = [int(cum[hi] - cum[lo]) for lo, hi in zip(z_lo, z_hi)]
    two = np.convolve(hist, np.ones(2, dtype=np.int64), mode="full")
    cap_own = int(1.5 * max(counts)) + 4096
    cap_ghost = int(2.0 * int(two.max())) + 4096
    return dict(z_lo=z_lo, z_hi=z_hi, counts=counts, cap_own=cap_own, cap_ghost=cap_ghost, bz=bz)
