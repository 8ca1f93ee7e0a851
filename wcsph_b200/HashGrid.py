"""HashGrid -- drop-in for the reference's HashGrid.py.

Same constructor signature (HashGrid.py:10), attributes (`searchR`, `gridR`, `invGridR`,
`maxInGrid`, `maxNeighbour`, `blockSize`, `min_boundary`, `max_boundary`, `neighborCount`,
`neighbor`) and methods (`setup_grid_gpu`, `setup_grid_cpu`, `update_grid`).  The N x 64
bucket table and the NL x 2048 candidate table are NOT materialised (8 KiB/particle,
SURVEY.md fact 9): update_grid() cell-sorts the liquids and builds compact in-range lists
in csrc/grid.cu; `neighborCount` is the reference-exact candidate count and `neighbor` a
lazy debug view.
"""
import ctypes as C

import numpy as np

from . import _lib
from .field import Field


class _NeighborView:
    """HashGrid.neighbor[i, k] as a lazy per-row view: in-range candidates of reference
    particle i as a multiset of reference indices (bucket-alias duplicates included)."""

    def __init__(self, grid):
        self._g = grid

    def row(self, i):
        pd = self._g.particle_data
        cap = 4096
        out = np.empty(cap, dtype=np.int32)
        n = C.c_int()
        _lib.check(_lib.load().wcsph_hashgrid_neighbors_of(pd._ctx, int(i), out.ctypes.data, cap, C.byref(n)))
        return out[: min(n.value, cap)].copy()

    def __getitem__(self, ik):
        i, k = ik
        return self.row(i)[k]


class HashGrid:
    def __init__(self, gridR, maxInGrid, maxNeighbour, particle_data):
        self.maxInGrid = maxInGrid
        self.maxNeighbour = maxNeighbour
        self.particle_data = particle_data
        self.invGridR = 1.0 / gridR
        self.gridR = gridR
        self.searchR = gridR * 2.0
        self.blockSize = np.ones(shape=(1, 3), dtype=np.int32)
        self.min_boundary = np.ones(shape=(1, 3), dtype=np.float32)
        self.max_boundary = np.ones(shape=(1, 3), dtype=np.float32)
        self.neighborCount = None
        self.neighbor = None

    def setup_grid_gpu(self):
        """HashGrid.py:34-40: the tables live in ParticleData's device arena."""
        self.neighborCount = Field(self.particle_data, "neighborCount")
        self.neighbor = _NeighborView(self)

    def setup_grid_cpu(self, maxboundarynp, minboundarynp):
        """HashGrid.py:44-54."""
        blocknp = np.ones(shape=(1, 3), dtype=np.int32)
        for i in range(3):
            # np.float32 difference, then float64 division: the numpy-1.x scalar promotion the reference ran under
            blocknp[0, i] = int(float(maxboundarynp[0, i] - minboundarynp[0, i]) / self.gridR + 1)
        self.max_boundary = np.array(maxboundarynp, dtype=np.float32)
        self.min_boundary = np.array(minboundarynp, dtype=np.float32)
        self.blockSize = blocknp
        dev = (C.c_int * 3)()
        _lib.check(_lib.load().wcsph_block_size(self.particle_data._ctx, C.byref(dev)))
        if tuple(dev) != tuple(int(x) for x in blocknp[0]):
            raise _lib.WcsphError("blockSize mismatch host %s device %s" % (blocknp[0], tuple(dev)))
        if self.particle_data.verbose:
            print("serach grid szie:", int(blocknp[0, 0] * blocknp[0, 1] * blocknp[0, 2]))

    def update_grid(self):
        """HashGrid.py:57-85."""
        _lib.check(_lib.load().wcsph_hashgrid_update_grid(self.particle_data._ctx))

    def status(self):
        """device status bits since the last call (HashGrid.py:73,103 overflow prints)."""
        f = C.c_uint32()
        _lib.check(_lib.load().wcsph_status(self.particle_data._ctx, C.byref(f)))
        return f.value
