"""ParticleData -- drop-in for the reference's ParticleData.py (container + constants + ingest).

Keeps the constructor `ParticleData(particleR)`, the counters, the physical constants with
the reference's spelling (`liqiudMass`), `add_liquid_point / add_solid_point / add_obj /
setup_data_gpu / setup_data_cpu` in the reference's call order (dfsph.py:66-82), and every
field attribute the solver scripts touch (ParticleData.py:33-74).  Fields are `Field`
shims over ONE torch uint8 CUDA tensor (the arena) that libwcsph_b200 sub-allocates.

`mc_grid` (ParticleData.py:29,177,184) is built on first access -- the reference allocates its 3M-vertex buffer and the
dense grid for every scene even though no step loop reads them (Q21).  Not built: `color`, `color_grad`, `pos_avr`, `G`
(anisotropic-kernel pre-pass, switched off in the reference's export_surface).
"""
import ctypes as C

import numpy as np

from . import _lib
from .HashGrid import HashGrid
from .constants import solver_params
from .field import Field, ScalarField
from .kernels.CubicKernel import CubicKernel

_VEC_FIELDS = ("pos", "normal", "vel_guess", "vel", "omega", "d_vel", "d_omega", "cg_r", "cg_dir", "cg_Ad", "cg_s",
               "d_ii", "dij_pj", "pos_star", "vel_star", "d_vel_pre")
_SCALAR_FIELDS = ("vel_max", "pressure", "rho", "adv_rho", "alpha_coff", "kappa", "kappa_v", "a_ii", "pressure_pre")
_ONE = ("avg_density_err", "cg_delta", "cg_delta_old", "cg_delta_zero", "rho_err", "deltaT", "vel_max0")


class _Gravity(tuple):
    """ti.Vector([0,-9.81,0]) stand-in: indexable, and `.x/.y/.z`."""
    x = property(lambda s: s[0])
    y = property(lambda s: s[1])
    z = property(lambda s: s[2])


class ParticleData:
    def __init__(self, particleR, solver="dfsph", constants=None, list_cap_liquid=0, list_cap_solid=0,
                 cull_scale=0.0, verbose=False, world_size=1, rank=0):
        self.particleR = particleR
        self.count = 0
        self.liquid_count = 0
        self.solid_count = 0
        self.verbose = verbose

        # ParticleData.py:18-22
        self.rho_L0 = 1000.0
        self.rho_S0 = self.rho_L0
        self.VL0 = particleR * particleR * particleR * 0.8 * 8.0
        self.VS0 = self.VL0
        self.liqiudMass = self.VL0 * self.rho_L0

        self.hash_grid = HashGrid(particleR * 2.0, 64, 2048, self)      # ParticleData.py:27
        self.kernel_c = CubicKernel(self.hash_grid.searchR)            # ParticleData.py:31
        self._mc_grid = None

        # ParticleData.py:61-65, :80-81, :85-87
        self.gravity = _Gravity((0.0, -9.81, 0.0))
        self.dim_coff = 10.0
        self.viscosity = 10.0
        self.viscosity_b = 10.0
        self.viscosity_err = 0.05
        self.tension_coff = 0.0
        self.tension_coff_b = 0.0
        self.viscosity_omega = 0.1
        self.vorticity_coff = 0.01
        self.vorticity_init = 0.5

        self.point_list = []
        self._chunks = []            # vectorised bulk ingest (arrays), in insertion order
        self.maxboundarynp = np.ones(shape=(1, 3), dtype=np.float32)
        self.minboundarynp = np.ones(shape=(1, 3), dtype=np.float32)
        for j in range(3):
            self.maxboundarynp[0, j] = -10000.0
            self.minboundarynp[0, j] = 10000.0

        self.solver = solver
        self._constants = constants      # solver module namespace (sesph/pcisph/iisph keep their own)
        self._list_caps = (list_cap_liquid, list_cap_solid)
        self._cull_scale = cull_scale
        self._ctx = None
        self._arena = None
        self._solids_started = False
        # z-slab decomposition over the GPUs of one box: one process per GPU, torch.distributed
        # (already initialised by the caller) carries the NCCL id, the library does the exchanges
        self.world_size, self.rank = int(world_size), int(rank)

    # ---- ingest (ParticleData.py:100-138) ------------------------------------------------
    def _grow_bbox(self, pts):
        p32 = pts.astype(np.float32)
        self.maxboundarynp[0] = np.maximum(self.maxboundarynp[0], p32.max(axis=0))
        self.minboundarynp[0] = np.minimum(self.minboundarynp[0], p32.min(axis=0))

    def add_liquid_point(self, point):
        if self._solids_started:
            raise ValueError("liquid points must be added before solid points (dfsph.py:258 index contract)")
        self.point_list.append(point)
        for j in range(3):
            self.maxboundarynp[0, j] = max(self.maxboundarynp[0, j], point[j])
            self.minboundarynp[0, j] = min(self.minboundarynp[0, j], point[j])
        self._chunks.append(np.asarray([point], dtype=np.float64))
        self.count += 1
        self.liquid_count += 1

    def add_solid_point(self, point):
        self._solids_started = True
        self.point_list.append(point)
        for j in range(3):
            self.maxboundarynp[0, j] = max(self.maxboundarynp[0, j], point[j])
            self.minboundarynp[0, j] = min(self.minboundarynp[0, j], point[j])
        self._chunks.append(np.asarray([point], dtype=np.float64))
        self.count += 1
        self.solid_count += 1

    def add_liquid_points(self, pts):
        """vectorised add_liquid_point for 1M+ scenes (same order, same bbox rule)."""
        if self._solids_started:
            raise ValueError("liquid points must be added before solid points")
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        if len(pts) == 0:
            return
        self._chunks.append(pts)
        self._grow_bbox(pts)
        self.count += len(pts)
        self.liquid_count += len(pts)

    def add_solid_points(self, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        if len(pts) == 0:
            return
        self._solids_started = True
        self._chunks.append(pts)
        self._grow_bbox(pts)
        self.count += len(pts)
        self.solid_count += len(pts)

    def add_obj(self, filename):
        from .scenes import load_boundary
        self.add_solid_points(load_boundary(filename))

    # ---- allocation / upload (ParticleData.py:142-185) --------------------------------------
    def _namespace(self):
        """constants the kernels bake in: the solver module's own, else ParticleData's (dfsph)."""
        if self._constants is not None:
            ns = dict(self._constants)
        else:
            ns = {k: getattr(self, k) for k in ("rho_L0", "rho_S0", "VL0", "VS0", "liqiudMass", "dim_coff", "viscosity",
                                                "viscosity_b", "viscosity_err", "tension_coff", "tension_coff_b",
                                                "viscosity_omega", "vorticity_coff", "vorticity_init")}
            ns["gravity"] = tuple(self.gravity)
            ns["searchR"] = self.hash_grid.searchR
            ns["kernel_style"] = 0
        return ns

    def params(self):
        return solver_params(self._namespace())

    def setup_data_gpu(self):
        import torch
        if not torch.cuda.is_available():
            raise _lib.WcsphError("wcsph_b200 needs a CUDA device; there is no CPU fallback")
        L = _lib.load()
        d = _lib.Desc()
        d.abi_version = _lib.ABI_VERSION
        d.solver = _lib.SOLVER_ID[self.solver]
        d.count, d.liquid_count = self.count, self.liquid_count
        d.hash_gridR = self.hash_grid.gridR
        d.max_in_grid, d.max_neighbour = self.hash_grid.maxInGrid, self.hash_grid.maxNeighbour
        d.list_cap_liquid, d.list_cap_solid = self._list_caps
        d.cull_scale = self._cull_scale
        for k in range(3):
            d.min_boundary[k] = float(self.minboundarynp[0, k])
            d.max_boundary[k] = float(self.maxboundarynp[0, k])
        d.params = self.params()
        if self.world_size > 1:
            from .partition import z_slabs
            pts = np.concatenate(self._chunks, axis=0)[: self.liquid_count]
            plan = z_slabs(pts, self.minboundarynp[0], self.maxboundarynp[0], self.hash_grid.gridR, self.world_size)
            d.world_size, d.rank = self.world_size, self.rank
            d.z_lo, d.z_hi = plan["z_lo"][self.rank], plan["z_hi"][self.rank]
            d.cap_own, d.cap_ghost = plan["cap_own"], plan["cap_ghost"]
            self.slab_plan = plan
        nbytes = L.wcsph_arena_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.WcsphError(L.wcsph_last_error().decode())
        self._arena = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        self._stream = torch.cuda.current_stream()
        ctx = C.c_void_p()
        _lib.check(L.wcsph_create(C.byref(d), C.c_void_p(self._arena.data_ptr()), nbytes,
                                  C.c_void_p(self._stream.cuda_stream), C.byref(ctx)))
        self._ctx = ctx
        self._desc = d
        if self.world_size > 1:
            import torch.distributed as dist
            idt = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                _lib.check(L.wcsph_comm_unique_id(buf, None))
                idt = torch.tensor(list(buf), dtype=torch.uint8)
            dev = idt.cuda() if dist.get_backend() == "nccl" else idt
            dist.broadcast(dev, src=0)
            raw = bytes(dev.cpu().tolist())
            _lib.check(L.wcsph_comm_init(ctx, raw, None))
            self.p2p_scalars = self._open_mailboxes(L, ctx, dist)
        for n in _VEC_FIELDS + _SCALAR_FIELDS + ("cg_Minv",):
            setattr(self, n, Field(self, n))
        for n in _ONE:
            setattr(self, n, ScalarField(self, n))
        self.hash_grid.setup_grid_gpu()

    def _open_mailboxes(self, L, ctx, dist):
        """peer mailboxes for the latency-bound exchanges of a step (include/wcsph_b200.h: wcsph_comm_mailbox_*): every rank exports
        the IPC handle of its mailbox, the handles are all-gathered, every rank maps its peers'.  Ranks that cannot map each other (no
        peer access, several nodes, a gloo group) stay on the NCCL calls -- decided collectively, so that all ranks take the same path."""
        import os
        import torch
        ok = 0
        h = (C.c_ubyte * 64)()
        if dist.get_backend() == "nccl" and os.environ.get("WCSPH_P2P_SCALARS", "1") != "0":
            ok = 1 if L.wcsph_comm_mailbox_handle(ctx, h) == 0 else 0
        mine = torch.tensor([ok] + list(h), dtype=torch.uint8)
        mine = mine.cuda() if dist.get_backend() == "nccl" else mine
        every = [torch.zeros_like(mine) for _ in range(self.world_size)]
        dist.all_gather(every, mine)
        every = [t.cpu() for t in every]
        opened = 0
        if all(int(t[0]) == 1 for t in every):
            blob = bytes(b for t in every for b in t[1:].tolist())
            opened = 1 if L.wcsph_comm_mailbox_open(ctx, blob) == 0 else 0
        flag = torch.tensor([opened], dtype=torch.int32)
        flag = flag.cuda() if dist.get_backend() == "nccl" else flag
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use = int(flag.item()) == 1
        if opened and not use:
            _lib.check(L.wcsph_set_option(ctx, b"p2p_scalars", 0))
        if self.verbose and self.rank == 0:
            print("peer mailboxes:", "on" if use else "off (NCCL scalars)")
        return use

    def setup_data_cpu(self):
        pts = np.concatenate(self._chunks, axis=0) if self._chunks else np.zeros((0, 3))
        pos32 = np.ascontiguousarray(pts, dtype=np.float32)          # ParticleData.py:182
        _lib.check(_lib.load().wcsph_upload_pos(self._ctx, pos32.ctypes.data))
        self.hash_grid.setup_grid_cpu(self.maxboundarynp, self.minboundarynp)
        if self.verbose:
            print("liqiud particle num:", self.liquid_count, "solid particle num:", self.solid_count)

    def update_params(self):
        """re-bake constants after the host changed one (the reference would re-JIT)."""
        p = self.params()
        _lib.check(_lib.load().wcsph_set_params(self._ctx, C.byref(p)))

    def call(self, fn, *args):
        f = getattr(_lib.load(), "wcsph_" + fn)
        _lib.check(f(self._ctx, *args))

    def iters(self):
        out = (C.c_int * 3)()
        _lib.check(_lib.load().wcsph_iters(self._ctx, C.byref(out)))
        return tuple(out)

    def launch_count(self, reset=False):
        return int(_lib.load().wcsph_launch_count(self._ctx, 1 if reset else 0))

    def sync(self):
        _lib.check(_lib.load().wcsph_sync(self._ctx))

    def check(self):
        """synchronise and RAISE if the device dropped pairs (compact-list stride, alias table, 64-slot bucket, far
        migration): the reference prints "exceed grid" / "exceed neighbor" (HashGrid.py:73,103) and carries on; here a run
        with missing pairs does not carry on silently.  `hash_grid.status()` reads and acknowledges the bits."""
        _lib.check(_lib.load().wcsph_check(self._ctx))

    def reupload(self, pos):
        """positions of EVERY particle in reference order -> fresh device state: on z-slab ranks this re-partitions the
        liquids by cell layer (a restored checkpoint may place them in other slabs than t = 0 did)."""
        pos32 = np.ascontiguousarray(pos, dtype=np.float32)
        if pos32.shape != (self.count, 3):
            raise ValueError("reupload wants %s positions, got %s" % ((self.count, 3), pos32.shape))
        _lib.check(_lib.load().wcsph_upload_pos(self._ctx, pos32.ctypes.data))

    # ---- surface reconstruction, built lazily (Q21) ---------------------------------
    @property
    def mc_grid(self):
        if self._mc_grid is None:
            from .MarchingCubeGrid import MCGrid
            g = MCGrid(self.particleR, 4, 512, self)                                     # ParticleData.py:29
            g.setup_grid_gpu(self.maxboundarynp + g.searchR, self.minboundarynp - g.searchR)   # ParticleData.py:177
            g.setup_grid_cpu(self.maxboundarynp + g.searchR, self.minboundarynp - g.searchR)   # ParticleData.py:184
            self._mc_grid = g
        return self._mc_grid

    # ---- the anisotropic-kernel pre-pass of the surface reconstruction (ParticleData.py:187-317) ------------
    # The reference keeps these kernels but has their call sites commented out (MarchingCubeGrid.py:148-149, dfsph.py:639).
    # color / color_grad / pos_avr / G are allocated on first use (Q21), live in the device's cell-sorted slot order and are
    # exposed in REFERENCE order through .to_numpy() like every other field.
    def _slot_field(self, name, tensor_fn, ncomp):
        pd = self

        class _SlotField:
            """reference-order view of a slot-ordered device tensor (stride 1, 4 or 12 floats per particle)"""

            def to_torch(self):
                return tensor_fn()

            def to_numpy(self):
                import torch
                t = tensor_fn()
                pd.sync()
                sid_ptr = C.c_void_p()
                _lib.check(_lib.load().wcsph_sorted_id_device(pd._ctx, C.byref(sid_ptr)))
                n = pd.liquid_count
                host = t[:n].cpu().numpy()
                # sorted slot -> reference index: an int32 table inside the arena
                off = sid_ptr.value - pd._arena.data_ptr()
                sid = pd._arena[off: off + n * 4].view(torch.int32).cpu().numpy()
                if ncomp == 1:
                    out = np.empty(n, np.float32); out[sid] = host
                elif ncomp == 3:
                    out = np.empty((n, 3), np.float32); out[sid] = host[:, :3]
                else:
                    out = np.empty((n, 3, 3), np.float32); out[sid] = host.reshape(n, 3, 4)[:, :, :3]
                return out
        return _SlotField()

    def _color_buffers(self):
        import torch
        if getattr(self, "_color_t", None) is None:
            n = max(self.liquid_count, 1)
            self._color_t = torch.zeros(n, dtype=torch.float32, device="cuda")
            self._color_grad_t = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        return self._color_t, self._color_grad_t

    def _aniso_workspace(self):
        import torch
        if getattr(self, "_aniso_work", None) is None:
            nbytes = _lib.load().wcsph_pd_aniso_workspace_bytes(self._ctx)
            self._aniso_work = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        return self._aniso_work

    def _aniso_buffers(self, need=False):
        import torch
        if getattr(self, "_pos_avr_t", None) is None:
            if need:
                raise _lib.WcsphError("call particle_data.cal_anistropic_kernel() first")
            n = max(self.liquid_count, 1)
            self._pos_avr_t = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
            self._G_t = torch.zeros((n, 12), dtype=torch.float32, device="cuda")
        return self._pos_avr_t, self._G_t

    def compute_color_map(self):
        """ParticleData.py:187-218: color[i], color_grad[i] from the state the last step left (rho, neighbour lists)."""
        import torch
        if self.world_size > 1:
            raise _lib.WcsphError("compute_color_map runs on a single-GPU context")
        c, g = self._color_buffers()
        w = self._aniso_workspace()
        torch.cuda.current_stream().synchronize()
        _lib.check(_lib.load().wcsph_pd_compute_color_map(self._ctx, C.c_void_p(w.data_ptr()), w.numel(), C.c_void_p(c.data_ptr()), C.c_void_p(g.data_ptr())))

    def cal_anistropic_kernel(self):
        """ParticleData.py:220-285: pos_avr[i] and the anisotropy matrix G[i] (3x3 SVD -> symmetric eigen-decomposition)."""
        import torch
        if self.world_size > 1:
            raise _lib.WcsphError("cal_anistropic_kernel runs on a single-GPU context")
        pa, G = self._aniso_buffers()
        w = self._aniso_workspace()
        torch.cuda.current_stream().synchronize()
        _lib.check(_lib.load().wcsph_pd_cal_anistropic_kernel(self._ctx, C.c_float(self.mc_grid.searchR), C.c_void_p(w.data_ptr()),
                                                              w.numel(), C.c_void_p(pa.data_ptr()), C.c_void_p(G.data_ptr())))

    @property
    def color(self):
        return self._slot_field("color", lambda: self._color_buffers()[0], 1)

    @property
    def color_grad(self):
        return self._slot_field("color_grad", lambda: self._color_buffers()[1], 3)

    @property
    def pos_avr(self):
        return self._slot_field("pos_avr", lambda: self._aniso_buffers(need=True)[0], 3)

    @property
    def G(self):
        return self._slot_field("G", lambda: self._aniso_buffers(need=True)[1], 9)

    def export_kernel(self, filename="out/test.obj"):
        """ParticleData.py:302-311: `v x y z r g b gx gy gz` with color_grad as the colour."""
        import os
        pos = self.pos.to_numpy()
        color = self.color_grad.to_numpy()
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        with open(filename, "w") as fo:
            for i in range(self.liquid_count):
                fo.write("v %f %f %f %f %f %f %f %f %f\n" % (pos[i, 0], pos[i, 1], pos[i, 2], color[i, 0] * 512.0, color[i, 1] * 512.0,
                                                             color[i, 2] * 512.0, color[i, 0], color[i, 1], color[i, 2]))
        return filename

    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load().wcsph_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass
