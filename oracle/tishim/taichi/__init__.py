"""Serial Taichi-semantics shim -- TEST INFRASTRUCTURE ONLY.

A stand-in for the `taichi` package (absent from this image) that lets the
UNMODIFIED sources under /root/reference execute in plain CPython, one logical thread, loop
indices ascending.  It exists for one purpose: `tests/golden/make_ref_exec.py` runs the
reference's own kernels on small scenes under it and writes the per-kernel / per-step goldens
(`tests/golden/ref_exec_*.npz`) that pin `oracle/wcsph_oracle.c` (CPU tests) and the CUDA path
(GPU tests).  Nothing under wcsph_b200/ imports it; it cannot travel to the product.

It implements the language subset the reference uses, with the semantics of SURVEY.md 2.5:

* default_fp = f32, default_ip = i32.  Values that live in Taichi scope are `numpy.float32` /
  `numpy.int32` scalars; Python-scope values (module constants, `self.attr`, literals) keep
  Python semantics (float64 / unbounded int) until they meet a Taichi value, then narrow --
  which is how Taichi's AST transformer folds `self.m_l*self.h3` in Python before the product
  meets an Expr.
* i32 arithmetic wraps, `/` is true division in f32, `%` / `//` floor, `ti.cast(f32->i32)`
  truncates toward zero; a local keeps the type of its first assignment.
* every top-level `for` of a kernel runs serially, ascending, one after the other;
  `ti.atomic_add` returns the old value; `x += ..` on fields is a plain read-modify-write.
* fields are numpy arrays; `field[i]` in Python scope returns Python scalars, `to_numpy()` /
  `from_numpy()` copy.
* out-of-bounds field access (undefined in Taichi with debug=False; the reference has three:
  Q7 `pcisph.py:234`, Q12 `dfsph.py:324`, Q15 `dfsph.py:563`) READS AS ZERO, WRITES ARE DROPPED,
  and every event is counted in `oob_log` so the golden generator reports them instead of hiding them.

The per-function source rewriting that routes operators here lives in `../loader.py`.
"""
import builtins as _bi
import math as _math

import numpy as _np

F32 = _np.float32
I32 = _np.int32
_np.seterr(all="ignore")

# ----------------------------------------------------------------------------- dtypes / axes
f32 = "f32"
i32 = "i32"
f64 = "f64"
i64 = "i64"
u8 = "u8"
gpu = "gpu"
cpu = "cpu"
cuda = "cuda"


class _Axes(tuple):
    pass


i = _Axes((0,))
j = _Axes((1,))
k = _Axes((2,))
ij = _Axes((0, 1))
ijk = _Axes((0, 1, 2))

_NPD = {f32: _np.float32, i32: _np.int32, f64: _np.float64, i64: _np.int64, u8: _np.uint8}

# ----------------------------------------------------------------------------- scope state
_state = {"depth": 0, "kernel": None}
oob_log = {}          # (kernel name, field label, 'r'|'w') -> count
trace_hook = [None]   # callable(kernel_name, owner, 'pre'|'post') around every top-level kernel
launch_log = []
_rng = [_np.random.default_rng(20240607)]


def in_kernel():
    return _state["depth"] > 0


def _oob(field, kind):
    key = (_state["kernel"], field.label or "?", kind)
    oob_log[key] = oob_log.get(key, 0) + 1


# ----------------------------------------------------------------------------- scalar rules
_PYF = (float, _np.float64)
_PYI = (int, bool, _np.int64, _np.bool_)


def _kind(x):
    t = type(x)
    if t is F32:
        return 2
    if t is I32:
        return 1
    if t in (float, int, bool) or isinstance(x, (_np.float64, _np.int64, _np.bool_)):
        return 0
    return 3  # Matrix or foreign object


def _isfloat(x):
    return isinstance(x, (float, _np.floating))


def c(x):
    """value entering Taichi scope (first assignment of a local, element of ti.Vector([...]))."""
    t = type(x)
    if t is F32 or t is I32:
        return x
    if t is Matrix:
        return x.copy_ti()
    if isinstance(x, (bool, _np.bool_)):
        return I32(int(x))
    if isinstance(x, (int, _np.integer)):
        return I32(_wrap(int(x)))
    if isinstance(x, (float, _np.floating)):
        return F32(x)
    return x


def _wrap(v):
    v &= 0xFFFFFFFF
    return v - 0x100000000 if v >= 0x80000000 else v


def cast_like(old, new):
    """assignment to an existing local: the variable keeps its declared type."""
    t = type(old)
    if t is F32:
        if type(new) is Matrix:
            return new.copy_ti()
        return F32(new)
    if t is I32:
        if type(new) is Matrix:
            return new.copy_ti()
        return to_i32(new)
    if t is Matrix and type(new) is Matrix and len(old.e) == len(new.e):
        return Matrix._raw([cast_like(a, b) for a, b in zip(old.e, new.e)], new.n, new.m)
    return c(new)


def to_i32(x):
    if type(x) is I32:
        return x
    if isinstance(x, (float, _np.floating)):
        if x != x or x in (_math.inf, -_math.inf):
            return I32(-2147483648)
        return I32(_wrap(int(x)))           # trunc toward zero
    return I32(_wrap(int(x)))


def _arith(a, b, fop, iop, pop):
    """binary op on scalars with Taichi promotion (python op python stays python)."""
    ka, kb = _kind(a), _kind(b)
    if ka == 0 and kb == 0:
        return pop(a, b)
    if ka == 2 or kb == 2 or _isfloat(a) or _isfloat(b):
        return fop(F32(a), F32(b))
    return iop(to_i32(a), to_i32(b))


def _mat_bin(a, b, fn):
    if type(a) is Matrix:
        if type(b) is Matrix:
            assert a.n == b.n and a.m == b.m, "matrix shape mismatch"
            return Matrix._raw([fn(x, y) for x, y in zip(a.e, b.e)], a.n, a.m)
        return Matrix._raw([fn(x, b) for x in a.e], a.n, a.m)
    return Matrix._raw([fn(a, y) for y in b.e], b.n, b.m)


def add(a, b):
    if type(a) is F32 and type(b) is F32:
        return a + b
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, add)
    return _arith(a, b, lambda x, y: x + y, lambda x, y: x + y, lambda x, y: x + y)


def sub(a, b):
    if type(a) is F32 and type(b) is F32:
        return a - b
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, sub)
    return _arith(a, b, lambda x, y: x - y, lambda x, y: x - y, lambda x, y: x - y)


def mul(a, b):
    if type(a) is F32 and type(b) is F32:
        return a * b
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, mul)
    return _arith(a, b, lambda x, y: x * y, lambda x, y: x * y, lambda x, y: x * y)


def div(a, b):
    """true division: ints are cast to f32 first (default_fp)."""
    if type(a) is F32 and type(b) is F32:
        return a / b
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, div)
    if _kind(a) == 0 and _kind(b) == 0:
        return a / b
    return F32(a) / F32(b)


def floordiv(a, b):
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, floordiv)
    return _arith(a, b, lambda x, y: F32(_np.floor(x / y)), lambda x, y: x // y, lambda x, y: x // y)


def mod(a, b):
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, mod)
    return _arith(a, b, lambda x, y: F32(_np.mod(x, y)), lambda x, y: x % y, lambda x, y: x % y)


def power(a, b):
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, power)
    return _arith(a, b, lambda x, y: F32(_np.power(x, y)), lambda x, y: x ** y, lambda x, y: x ** y)


def _intop(name, pop):
    def op(a, b):
        if type(a) is Matrix or type(b) is Matrix:
            return _mat_bin(a, b, op)
        if _kind(a) == 0 and _kind(b) == 0:
            return pop(a, b)
        return pop(to_i32(a), to_i32(b))
    op.__name__ = name
    return op


bxor = _intop("bxor", lambda x, y: x ^ y)
band = _intop("band", lambda x, y: x & y)
bor = _intop("bor", lambda x, y: x | y)
shl = _intop("shl", lambda x, y: x << y)
shr = _intop("shr", lambda x, y: x >> y)


def neg(a):
    return -a


import operator as _op
_HOST_OPS = {"add": _op.add, "sub": _op.sub, "mul": _op.mul, "div": _op.truediv, "floordiv": _op.floordiv, "mod": _op.mod,
             "power": _op.pow, "bxor": _op.xor, "band": _op.and_, "bor": _op.or_, "shl": _op.lshift, "shr": _op.rshift}


def host_op(name, a, b):
    """Python-scope binary operator with NumPy 1.x scalar promotion (see loader._Host): a float32 SCALAR meeting a
    Python float is widened to float64 first; every other pairing (arrays, ints, Matrix, ...) is the plain operator."""
    ta, tb = type(a), type(b)
    if ta is F32 and (tb is float or tb is _np.float64):
        a = float(a)
    elif tb is F32 and (ta is float or ta is _np.float64):
        b = float(b)
    return _HOST_OPS[name](a, b)


# builtins as Taichi sees them in kernel scope
def ti_int(x):
    if _kind(x) == 0:
        return int(x)
    if type(x) is Matrix:
        return x._map(to_i32)
    return to_i32(x)


def ti_float(x):
    if _kind(x) == 0:
        return float(x)
    if type(x) is Matrix:
        return x._map(F32)
    return F32(x)


def _sel(a, b, take_a_py, fsel, isel):
    if type(a) is Matrix or type(b) is Matrix:
        return _mat_bin(a, b, lambda x, y: _sel(x, y, take_a_py, fsel, isel))
    return _arith(a, b, fsel, isel, take_a_py)


def ti_max(*a):
    r = a[0]
    for x in a[1:]:
        r = _sel(r, x, lambda p, q: p if p >= q else q, lambda p, q: p if p >= q or q != q else q,
                 lambda p, q: p if p >= q else q)
    return r


def ti_min(*a):
    r = a[0]
    for x in a[1:]:
        r = _sel(r, x, lambda p, q: p if p <= q else q, lambda p, q: p if p <= q or q != q else q,
                 lambda p, q: p if p <= q else q)
    return r


def ti_abs(x):
    if type(x) is Matrix:
        return x._map(ti_abs)
    return _bi.abs(x)


def ti_pow(a, b):
    return power(a, b)


def ti_range(*a):
    if in_kernel():
        return (I32(v) for v in range(*[int(x) for x in a]))
    return range(*a)


max = ti_max   # ti.max / ti.min / ti.abs
min = ti_min
abs = ti_abs


def _unary(pyfn, npfn):
    def f(x):
        if type(x) is Matrix:
            return x._map(f)
        if _kind(x) == 0:
            return pyfn(x)
        return F32(npfn(F32(x)))
    return f


sqrt = _unary(_math.sqrt, _np.sqrt)
sin = _unary(_math.sin, _np.sin)
cos = _unary(_math.cos, _np.cos)
tan = _unary(_math.tan, _np.tan)
asin = _unary(_math.asin, _np.arcsin)
acos = _unary(_math.acos, _np.arccos)
exp = _unary(_math.exp, _np.exp)
log = _unary(_math.log, _np.log)
floor = _unary(_math.floor, _np.floor)
ceil = _unary(_math.ceil, _np.ceil)


def cast(x, dt):
    if type(x) is Matrix:
        return x._map(lambda v: cast(v, dt))
    if dt == i32:
        return to_i32(x)
    if dt == f32:
        return F32(x)
    raise NotImplementedError(dt)


def random(dt=f32):
    return F32(_rng[0].random(dtype=_np.float32))


def seed(s):
    _rng[0] = _np.random.default_rng(s)


def static(x):
    return x


def template():
    return "template"


def ndrange(*dims):
    import itertools
    rs = []
    for d in dims:
        if isinstance(d, (tuple, list)):
            rs.append(range(int(d[0]), int(d[1])))
        else:
            rs.append(range(int(d)))
    if len(rs) == 1:
        return (I32(v) for v in rs[0])
    return (tuple(I32(v) for v in t) for t in itertools.product(*rs))


def atomic_add_at(container, index, v):
    old = container[index]
    container[index] = add(old, v)
    return old


# ----------------------------------------------------------------------------- Matrix / Vector
_XYZW = {"x": 0, "y": 1, "z": 2, "w": 3}


class Matrix:
    """n x m small matrix (a Vector is n x 1), entries row-major in a Python list."""
    __slots__ = ("e", "n", "m")

    def __init__(self, rows, dt=None):
        if rows and isinstance(rows[0], (list, tuple)):
            n, m = len(rows), len(rows[0])
            e = [v for r in rows for v in r]
        else:
            n, m = len(rows), 1
            e = list(rows)
        if in_kernel():
            e = [c(v) for v in e]
        object.__setattr__(self, "e", e)
        object.__setattr__(self, "n", n)
        object.__setattr__(self, "m", m)

    @staticmethod
    def _raw(e, n, m):
        r = Matrix.__new__(Matrix)
        object.__setattr__(r, "e", e)
        object.__setattr__(r, "n", n)
        object.__setattr__(r, "m", m)
        return r

    def copy_ti(self):
        return Matrix._raw([c(v) for v in self.e], self.n, self.m)

    def _map(self, fn):
        return Matrix._raw([fn(v) for v in self.e], self.n, self.m)

    # fields --------------------------------------------------------------
    @staticmethod
    def field(n, m=None, dtype=f32, shape=None, **kw):
        if isinstance(m, str):
            dtype, m = m, None
        return Field(dtype, (n,) if m is None else (n, m), shape)

    # element access -------------------------------------------------------
    def _idx(self, ix):
        if isinstance(ix, tuple):
            return int(ix[0]) * self.m + int(ix[1])
        return int(ix)

    def __getitem__(self, ix):
        return self.e[self._idx(ix)]

    def __setitem__(self, ix, v):
        k_ = self._idx(ix)
        self.e[k_] = cast_like(self.e[k_], v) if in_kernel() else v

    def __call__(self, a, b=0):
        return self.e[a * self.m + b]

    def __getattr__(self, name):
        if name in _XYZW:
            return self.e[_XYZW[name]]
        raise AttributeError(name)

    def __setattr__(self, name, v):
        if name in _XYZW:
            self[_XYZW[name]] = v
        else:
            raise AttributeError(name)

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter(self.e)

    # arithmetic (also reachable without the loader's rewriting, e.g. from Python scope) -------
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return mul(self, o)
    def __rmul__(self, o): return mul(o, self)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __neg__(self): return self._map(lambda v: -v)
    def __matmul__(self, o): return matmul(self, o)

    # Taichi's matrix methods, same evaluation order as taichi/lang/matrix.py ------------------
    def sum(self):
        r = self.e[0]
        for v in self.e[1:]:
            r = add(r, v)
        return r

    def norm_sqr(self):
        return mul(self, self).sum()

    def norm(self, eps=0):
        return sqrt(add(self.norm_sqr(), eps)) if eps else sqrt(self.norm_sqr())

    def normalized(self, eps=0):
        invlen = div(1.0, add(self.norm(), eps)) if eps else div(1.0, self.norm())
        return mul(invlen, self)

    def dot(self, o):
        return mul(self, o).sum()

    def cross(self, o):
        a, b = self.e, o.e
        return Matrix._raw([sub(mul(a[1], b[2]), mul(a[2], b[1])),
                            sub(mul(a[2], b[0]), mul(a[0], b[2])),
                            sub(mul(a[0], b[1]), mul(a[1], b[0]))], 3, 1)

    def outer_product(self, o):
        return Matrix._raw([mul(a, b) for a in self.e for b in o.e], self.n, o.n)

    def transpose(self):
        return Matrix._raw([self.e[r * self.m + cc] for cc in range(self.m) for r in range(self.n)], self.m, self.n)

    def cast(self, dt):
        return cast(self, dt)

    def max(self):
        r = self.e[0]
        for v in self.e[1:]:
            r = ti_max(r, v)
        return r

    def min(self):
        r = self.e[0]
        for v in self.e[1:]:
            r = ti_min(r, v)
        return r

    def determinant(self):
        a = self
        if self.n == 2:
            return sub(mul(a(0, 0), a(1, 1)), mul(a(0, 1), a(1, 0)))
        assert self.n == 3 and self.m == 3
        return add(sub(mul(a(0, 0), sub(mul(a(1, 1), a(2, 2)), mul(a(2, 1), a(1, 2)))),
                       mul(a(1, 0), sub(mul(a(0, 1), a(2, 2)), mul(a(2, 1), a(0, 2))))),
                   mul(a(2, 0), sub(mul(a(0, 1), a(1, 2)), mul(a(1, 1), a(0, 2)))))

    def inverse(self):
        """closed-form adjugate / determinant, the 3x3 branch of taichi Matrix.inverse()."""
        assert self.n == 3 and self.m == 3
        n = 3
        inv_det = div(1.0, self.determinant())

        def E(x, y):
            return self(x % n, y % n)
        out = [None] * 9
        for i_ in range(n):
            for j_ in range(n):
                out[j_ * 3 + i_] = mul(inv_det, sub(mul(E(i_ + 1, j_ + 1), E(i_ + 2, j_ + 2)),
                                                    mul(E(i_ + 2, j_ + 1), E(i_ + 1, j_ + 2))))
        return Matrix._raw(out, 3, 3)

    def to_numpy(self):
        a = _np.array([float(v) for v in self.e])
        return a.reshape(self.n, self.m) if self.m > 1 else a

    def __repr__(self):
        return "Matrix(%r, %dx%d)" % (self.e, self.n, self.m)


def matmul(a, b):
    assert a.m == b.n
    out = []
    for r in range(a.n):
        for cc in range(b.m):
            acc = mul(a.e[r * a.m], b.e[cc])
            for t in range(1, a.m):
                acc = add(acc, mul(a.e[r * a.m + t], b.e[t * b.m + cc]))
            out.append(acc)
    return Matrix._raw(out, a.n, b.m)


def Vector(vals, dt=None):
    return Matrix(list(vals))


Vector.field = lambda n, dtype=f32, shape=None, **kw: Field(dtype, (n,), shape)


def svd(A, dt=f32):
    """ti.svd(A) -> U, S (diagonal matrix), V with A = U S V^T, singular values descending.
    Taichi's 3x3 routine is the McAdams et al. fixed-sweep Jacobi method (third-party, un-vendored);
    here LAPACK in float64, narrowed to f32 -- same decomposition up to column signs and ~1e-6."""
    a = _np.array([float(v) for v in A.e], _np.float64).reshape(A.n, A.m)
    u, s, vh = _np.linalg.svd(a)
    mk = lambda M_: Matrix._raw([F32(v) for v in M_.reshape(-1)], 3, 3)
    return mk(u), mk(_np.diag(s)), mk(vh.T)


# ----------------------------------------------------------------------------- fields
class Field:
    """ti.field / ti.Vector.field / ti.Matrix.field over one numpy array (shape + element shape)."""

    def __init__(self, dtype, eshape=(), shape=None):
        self.dtype = dtype
        self.npd = _NPD[dtype]
        self.eshape = tuple(eshape)
        self.arr = None
        self.shape = None
        self.label = None
        self._scal = F32 if dtype == f32 else I32
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, (int, _np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.arr = _np.zeros(self.shape + self.eshape, self.npd)

    # indexing --------------------------------------------------------------
    def _ix(self, ix):
        if type(ix) is Matrix:
            ix = tuple(ix.e)
        elif not isinstance(ix, tuple):
            ix = (ix,)
        if len(ix) != len(self.shape):
            raise IndexError("field rank mismatch")
        out = []
        for v, s in zip(ix, self.shape):
            v = int(v)
            if v < 0 or v >= s:
                return None
            out.append(v)
        return tuple(out)

    def __getitem__(self, ix):
        t = self._ix(ix)
        if t is None:
            _oob(self, "r")
            if not in_kernel():
                raise IndexError("out-of-bounds field read from Python scope")
            if not self.eshape:
                return self._scal(0)
            n = self.eshape[0]
            m = self.eshape[1] if len(self.eshape) > 1 else 1
            return Matrix._raw([self._scal(0)] * (n * m), n, m)
        v = self.arr[t]
        if not self.eshape:
            return v if in_kernel() else v.item()
        if len(self.eshape) == 1:
            return Matrix._raw(list(v), self.eshape[0], 1) if in_kernel() else Matrix._raw(v.tolist(), self.eshape[0], 1)
        flat = v.reshape(-1)
        return Matrix._raw(list(flat) if in_kernel() else flat.tolist(), self.eshape[0], self.eshape[1])

    def __setitem__(self, ix, val):
        t = self._ix(ix)
        if t is None:
            _oob(self, "w")
            if not in_kernel():
                raise IndexError("out-of-bounds field write from Python scope")
            return
        if type(val) is Matrix:
            if self.dtype == i32:
                self.arr[t] = _np.array([to_i32(v) for v in val.e], _np.int32).reshape(self.eshape)
            else:
                self.arr[t] = _np.array(val.e, self.npd).reshape(self.eshape)
        elif self.dtype == i32:
            self.arr[t] = to_i32(val)
        else:
            self.arr[t] = val

    def __iter__(self):
        raise TypeError("struct-for over a field outside a kernel")

    def to_numpy(self):
        return self.arr.copy()

    def from_numpy(self, a):
        a = _np.asarray(a)
        assert a.shape == self.arr.shape, (a.shape, self.arr.shape)
        self.arr[...] = a.astype(self.npd)

    def fill(self, v):
        self.arr[...] = v


def field(dtype=f32, shape=None, **kw):
    return Field(dtype, (), shape)


def struct_iter(x):
    """top-level `for i in field` / `for i, j in field`: ascending, row-major."""
    if isinstance(x, Field):
        if len(x.shape) == 1:
            return (I32(v) for v in range(x.shape[0]))
        import itertools
        return (tuple(I32(v) for v in t) for t in itertools.product(*[range(s) for s in x.shape]))
    return iter(x)


class _SNode:
    def dense(self, axes, shape):
        return _Dense(axes, shape)


class _Dense:
    def __init__(self, axes, shape):
        self.shape = shape

    def place(self, *fields):
        for f_ in fields:
            f_._alloc(self.shape)

    def dense(self, axes, shape):
        raise NotImplementedError("nested dense")


root = _SNode()


# ----------------------------------------------------------------------------- decorators / runtime
def init(*a, **kw):
    return None


def data_oriented(cls):
    return cls


def _label_fields(owner):
    """name fields after the attribute that holds them (for oob_log), lazily."""
    d = getattr(owner, "__dict__", None)
    if d is None:
        return
    for k_, v in list(d.items()):
        if isinstance(v, Field) and v.label is None:
            v.label = k_


def kernel(fn):
    import functools
    import inspect
    ann = fn.__annotations__
    names = list(inspect.signature(fn).parameters)

    @functools.wraps(fn)
    def run(*args):
        args = list(args)
        for n_, a in enumerate(args):
            t = ann.get(names[n_])
            if t == i32:
                args[n_] = to_i32(a)
            elif t == f32:
                args[n_] = F32(a)
        owner = args[0] if names and names[0] == "self" else None
        top = _state["depth"] == 0
        if top:
            _state["kernel"] = fn.__qualname__
            if owner is not None:
                _label_fields(owner)
            g = fn.__globals__
            for k_, v in list(g.items()):
                if isinstance(v, Field) and v.label is None:
                    v.label = k_
                elif hasattr(v, "__dict__") and not isinstance(v, type) and type(v).__module__ not in ("builtins", "types"):
                    _label_fields(v)          # e.g. the module global `particle_data`
            if trace_hook[0] is not None:
                trace_hook[0](fn.__name__, owner, "pre")
        _state["depth"] += 1
        try:
            fn(*args)
        finally:
            _state["depth"] -= 1
        if top:
            launch_log.append(fn.__qualname__)
            if trace_hook[0] is not None:
                trace_hook[0](fn.__name__, owner, "post")
    run.__ti_kernel__ = True
    return run


def func(fn):
    import functools

    @functools.wraps(fn)
    def call(*args):
        return fn(*[a.copy_ti() if type(a) is Matrix else a for a in args])
    return call


class GUI:
    """ti.GUI stand-in: `running` turns False after `max_frames` calls of show()."""
    max_frames = 1
    on_show = None

    def __init__(self, name="", res=(512, 512), **kw):
        self.name = name
        self.res = res
        self.frames = 0
        self.image = None

    @property
    def running(self):
        return self.frames < GUI.max_frames

    def set_image(self, img):
        self.image = img

    def show(self, *a):
        self.frames += 1
        if GUI.on_show is not None:
            GUI.on_show(self)


def imwrite(img, path):
    return None
