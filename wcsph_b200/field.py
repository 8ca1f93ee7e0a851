"""Field shim: the slice of the Taichi field API the reference's host code uses
(`.to_numpy()` dfsph.py:113, `.from_numpy()` dfsph.py:129, `field[0]` dfsph.py:98, `.shape`),
backed by the device arena of libwcsph_b200.  Values cross in the reference's insertion
order; on the device the liquids are cell-sorted."""
import ctypes as C

import numpy as np

from . import _lib


class Field:
    def __init__(self, owner, name):
        self._o = owner           # ParticleData (owns the ctx)
        self.name = name

    def _info(self):
        n, nc, ii = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().wcsph_field_info(self._o._ctx, self.name.encode(), C.byref(n), C.byref(nc), C.byref(ii)))
        return n.value, nc.value, ii.value

    @property
    def shape(self):
        n, nc, _ = self._info()
        return (n,) if nc == 1 else ((n, 3) if nc == 3 else (n, 3, 3))

    def to_numpy(self, gather=True):
        """reference-order copy.  On a z-slab rank the library fills only the rows this rank owns
        (zeros elsewhere); gather=True sums the ranks' arrays so every rank returns the full field."""
        n, nc, is_int = self._info()
        out = np.zeros((n, nc) if nc > 1 else (n,), dtype=np.int32 if is_int else np.float32)
        _lib.check(_lib.load().wcsph_field_get(self._o._ctx, self.name.encode(), out.ctypes.data, out.nbytes))
        if gather and getattr(self._o, "world_size", 1) > 1:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(out)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.all_reduce(t)
            out = t.cpu().numpy()
        return out.reshape(n, 3, 3) if nc == 9 else out

    def from_numpy(self, arr):
        n, nc, is_int = self._info()
        a = np.ascontiguousarray(arr, dtype=np.float32).reshape(n, nc) if nc > 1 else np.ascontiguousarray(arr, dtype=np.float32).reshape(n)
        _lib.check(_lib.load().wcsph_field_set(self._o._ctx, self.name.encode(), a.ctypes.data, a.nbytes))

    def __getitem__(self, i):
        return self.to_numpy()[i]

    def to_torch(self):
        """zero-copy torch view in CURRENT cell-sorted order, shape (n, stride)."""
        import torch
        p, n, s = C.c_void_p(), C.c_int(), C.c_int()
        _lib.check(_lib.load().wcsph_field_device(self._o._ctx, self.name.encode(), C.byref(p), C.byref(n), C.byref(s)))
        arena = self._o._arena
        off = p.value - arena.data_ptr()
        dt = torch.int32 if self.name == "neighborCount" else torch.float32
        return arena[off: off + n.value * s.value * 4].view(dt).view(n.value, s.value)


class ScalarField:
    """1-element field (deltaT, avg_density_err, cg_delta, ...): `f[0]`, `.to_numpy()`, `.from_numpy()`."""

    def __init__(self, owner, name):
        self._o = owner
        self.name = name
        self.shape = (1,)

    def to_numpy(self):
        v = C.c_float()
        _lib.check(_lib.load().wcsph_scalar_get(self._o._ctx, self.name.encode(), C.byref(v)))
        return np.array([v.value], dtype=np.float32)

    def from_numpy(self, arr):
        _lib.check(_lib.load().wcsph_scalar_set(self._o._ctx, self.name.encode(), float(np.asarray(arr).reshape(-1)[0])))

    def __getitem__(self, i):
        if i != 0:
            raise IndexError(i)
        return self.to_numpy()[0]

    def __setitem__(self, i, v):
        if i != 0:
            raise IndexError(i)
        self.from_numpy([v])
