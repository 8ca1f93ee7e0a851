"""MCGrid -- the surface reconstruction that consumes the path's output.

Mirror of the reference's `MCGrid(particleR, maxInGrid, maxNeighbour, particle_data)` (MarchingCubeGrid.py:11-409):
`setup_grid_gpu / setup_grid_cpu(maxboundarynp, minboundarynp)`, `update_grid()`, `cal_surface_point()`,
`marching_cube()`, `export_mesh()`, `export_vertex()`, `export_surface(time)`, the attributes the reference exposes
(`gridR, invGridR, searchR, grid_num, isolevel, fps, frame, blocknp, maxboundarynp, minboundarynp`) and the fields
`surface_value`, `triangle`, `vertex_count` with `.to_numpy()`.  Each former @ti.kernel is one call into
libwcsph_b200 (`csrc/mc.cu`); the case tables the reference parses from MCData.txt are compiled into the library.

`ParticleData.mc_grid` builds one lazily with the reference's arguments (ParticleData.py:29,177,184).
`cal_surface_point_anistropic()` (:215-246) is built too; the reference keeps its two calls commented out of `export_surface`
(:148-149), `export_surface(time, anisotropic=True)` runs them.
No CPU fallback: a missing library raises.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

MAX_VERTEX = 3000000            # MarchingCubeGrid.py:8


class _DeviceArray:
    """`.to_numpy()` / `.to_torch()` view of a device tensor the grid owns."""

    def __init__(self, get):
        self._get = get

    def to_torch(self):
        return self._get()

    def to_numpy(self):
        return self._get().cpu().numpy()

    @property
    def shape(self):
        return tuple(self._get().shape)


class MCGrid:
    def __init__(self, particleR, maxInGrid, maxNeighbour, particle_data, max_vertex=MAX_VERTEX):
        self.fps, self.frame = 20.0, 0
        self.maxInGrid, self.maxNeighbour = maxInGrid, maxNeighbour
        self.particle_data = particle_data
        self.gridR = particleR * 0.9                    # MarchingCubeGrid.py:22-27
        self.invGridR = 1.0 / self.gridR
        self.searchR = self.gridR * 4.0
        self.grid_num = 0
        self.isolevel = 0.5
        self.max_vertex = int(max_vertex) - int(max_vertex) % 3
        self.maxboundarynp = np.ones(shape=(1, 3), dtype=np.float32)
        self.minboundarynp = np.ones(shape=(1, 3), dtype=np.float32)
        self.blocknp = np.ones(shape=(1, 3), dtype=np.int32)
        self.out_dir = "out"                            # the reference writes out/<frame>.obj, out/mc_<frame>.obj
        self._desc = None
        self._work = self._sv = self._tri = None
        self._nvert = 0
        self.surface_value = _DeviceArray(lambda: self._need(self._sv, "cal_surface_point"))
        self.triangle = _DeviceArray(lambda: self._need(self._tri, "marching_cube"))
        self.vertex_count = _DeviceArray(lambda: self._count_tensor())

    # ---- setup (MarchingCubeGrid.py:56-96) ----
    def setup_grid_gpu(self, maxboundarynp, minboundarynp):
        for k in range(3):
            self.maxboundarynp[0, k] = maxboundarynp[0, k]
            self.minboundarynp[0, k] = minboundarynp[0, k]
            self.blocknp[0, k] = int(float(self.maxboundarynp[0, k] - self.minboundarynp[0, k]) / self.gridR + 1)
        self.grid_num = int(self.blocknp[0, 0]) * int(self.blocknp[0, 1]) * int(self.blocknp[0, 2])
        d = _lib.McGrid()
        d.gridR, d.isolevel, d.max_in_grid = self.gridR, self.isolevel, int(self.maxInGrid)
        d.liqiudMass = float(self.particle_data.liqiudMass)
        for k in range(3):
            d.min_boundary[k] = float(self.minboundarynp[0, k])
            d.block[k] = int(self.blocknp[0, k])
        self._desc = d

    def setup_grid_cpu(self, maxboundarynp, minboundarynp):
        """the reference loads MCData.txt here (:80-94); the tables are part of the library, only the message stays."""
        if getattr(self.particle_data, "verbose", False):
            print("MC grid szie:", self.grid_num, "MC grid R:", self.gridR)

    # ---- device side ----
    def _need(self, t, producer):
        if t is None:
            raise _lib.WcsphError("MCGrid: call %s() first" % producer)
        return t

    def _count_tensor(self):
        import torch
        return torch.tensor([self._nvert], dtype=torch.int32)

    def _buffers(self):
        import torch
        pd = self.particle_data
        if pd._ctx is None or self._desc is None:
            raise _lib.WcsphError("MCGrid: setup_data_gpu() / setup_grid_gpu() first")
        if self._work is None:
            L = _lib.load()
            nbytes = L.wcsph_mc_workspace_bytes(C.byref(self._desc), int(pd.liquid_count))
            if nbytes == 0:
                raise _lib.WcsphError("MCGrid: grid of %d nodes is not addressable" % self.grid_num)
            self._work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            self._sv = torch.empty(self.grid_num, dtype=torch.float32, device="cuda")
        return C.c_void_p(self._work.data_ptr()), self._work.numel()

    def update_grid(self):
        """MarchingCubeGrid.py:160-179."""
        w, n = self._buffers()
        _lib.check(_lib.load().wcsph_mc_update_grid(self.particle_data._ctx, C.byref(self._desc), w, n))
        self._nvert = 0                                     # :164

    def cal_surface_point(self):
        """MarchingCubeGrid.py:183-209."""
        w, n = self._buffers()
        _lib.check(_lib.load().wcsph_mc_cal_surface_point(self.particle_data._ctx, C.byref(self._desc), w, n,
                                                          C.c_void_p(self._sv.data_ptr())))
        if getattr(self.particle_data, "world_size", 1) > 1:
            # z-slab ranks: every rank evaluated the contributions of its own liquids; the node field is their sum (like the
            # canvas, whose per-pixel keys are min-reduced).  Every rank then holds the whole field and polygonises it.
            import torch
            import torch.distributed as dist
            self.particle_data.sync()
            torch.cuda.current_stream().synchronize()
            dist.all_reduce(self._sv)
            torch.cuda.current_stream().synchronize()

    def cal_surface_point_anistropic(self):
        """MarchingCubeGrid.py:215-246: colour field with the anisotropic kernels of ParticleData.cal_anistropic_kernel()
        (call update_grid() and particle_data.cal_anistropic_kernel() first, like the commented lines :147-149 do)."""
        w, n = self._buffers()
        pa, G = self.particle_data._aniso_buffers(need=True)
        _lib.check(_lib.load().wcsph_mc_cal_surface_point_anistropic(self.particle_data._ctx, C.byref(self._desc), w, n,
                                                                     C.c_void_p(pa.data_ptr()), C.c_void_p(G.data_ptr()),
                                                                     C.c_void_p(self._sv.data_ptr())))

    def marching_cube(self, surface_value=None):
        """MarchingCubeGrid.py:262-352.  `surface_value` (numpy / torch, grid_num f32) overrides the field for this call."""
        import torch
        w, n = self._buffers()
        sv = self._sv
        if surface_value is not None:
            sv = torch.as_tensor(np.ascontiguousarray(surface_value, np.float32) if isinstance(surface_value, np.ndarray)
                                 else surface_value).to(device="cuda", dtype=torch.float32).contiguous()
            if sv.numel() != self.grid_num:
                raise _lib.WcsphError("MCGrid.marching_cube: surface_value has %d entries, grid has %d" % (sv.numel(), self.grid_num))
            torch.cuda.current_stream().synchronize()          # the upload ran on torch's stream, the library runs on the context's
        if self._tri is None:
            self._tri = torch.zeros((self.max_vertex, 3), dtype=torch.float32, device="cuda")
        cnt = C.c_int(0)
        _lib.check(_lib.load().wcsph_mc_marching_cube(self.particle_data._ctx, C.byref(self._desc), w, n, C.c_void_p(sv.data_ptr()),
                                                      C.c_void_p(self._tri.data_ptr()), self.max_vertex, C.byref(cnt)))
        self._nvert = int(cnt.value)
        if self._nvert > self.max_vertex and getattr(self.particle_data, "verbose", False):
            print("exceed max tri", self._nvert)            # :349
        return self._nvert

    # ---- export (MarchingCubeGrid.py:100-157) ----
    def mesh(self):
        """(vertices [n,3] f32, n = min(vertex_count, max_vertex)); triangle t is rows 3t..3t+2."""
        n = min(self._nvert, self.max_vertex)
        return self._need(self._tri, "marching_cube")[:n].cpu().numpy()

    def export_mesh(self):
        v = self.mesh()
        os.makedirs(self.out_dir, exist_ok=True)
        path = os.path.join(self.out_dir, "mc_" + str(self.frame) + ".obj")
        with open(path, "w") as fo:
            fo.write("".join("v %f %f %f\n" % (p[0], p[1], p[2]) for p in v))
            fo.write("".join("f %d %d %d\n" % (3 * t + 1, 3 * t + 2, 3 * t + 3) for t in range(len(v) // 3)))
        return path

    def export_vertex(self):
        """:100-114 (the debug_value column of the reference is never written by any kernel: exported as 0)."""
        iso = self.surface_value.to_numpy()
        idx = np.nonzero(iso > 0.0)[0]
        by, bz = int(self.blocknp[0, 1]), int(self.blocknp[0, 2])
        x = (idx // (by * bz)).astype(np.float64) * self.gridR + self.minboundarynp[0, 0]
        y = ((idx % (by * bz)) // bz).astype(np.float64) * self.gridR + self.minboundarynp[0, 1]
        z = (idx % bz).astype(np.float64) * self.gridR + self.minboundarynp[0, 2]
        os.makedirs(self.out_dir, exist_ok=True)
        path = os.path.join(self.out_dir, str(self.frame) + ".obj")
        with open(path, "w") as fo:
            fo.write("".join("v %f %f %f %f %f %f\n" % (x[k], y[k], z[k], iso[idx[k]], 0.0, 1.0) for k in range(len(idx))))
        return path

    def export_surface(self, time, anisotropic=False):
        """:137-157: one mesh per 1/fps of simulated time.  anisotropic=True takes the branch the reference has commented out
        (:147-149): particle_data.cal_anistropic_kernel() + cal_surface_point_anistropic()."""
        if int(time * self.fps) == self.frame:
            self.update_grid()
            if anisotropic:
                self.particle_data.cal_anistropic_kernel()
                self.cal_surface_point_anistropic()
            else:
                self.cal_surface_point()
            self.marching_cube()
            self.export_mesh()
            self.frame += 1
