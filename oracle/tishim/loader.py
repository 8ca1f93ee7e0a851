"""Load UNMODIFIED reference modules under the serial Taichi shim -- TEST INFRASTRUCTURE ONLY.

`RefLoader(ref_dir, consts)` installs an import hook that serves `HashGrid`, `ParticleData`,
`kernels.*`, `Canvas`, `MarchingCubeGrid`, `sesph`, ... straight from the reference checkout.  The
source text is the reference's; two mechanical things happen to its AST before it is compiled:

1. **module constants** named in `consts[module]` (e.g. `particleDimX`) get the given literal --
   the same substitution `tests/golden/make_fixtures.py` uses -- so that pure Python finishes a
   scene in minutes.  Nothing else at module level is touched.
2. inside every `@ti.kernel` / `@ti.func` body, operators and a handful of builtins are routed to
   the shim (`taichi/__init__.py`) so that arithmetic follows Taichi's typing rules instead of
   numpy's (int32*float32 -> f32, int/int -> f32, a local keeps its first type, struct-for over a
   field, `ti.atomic_add(field[i], v)` returning the old value).  Statement order, control flow and
   expression grouping are exactly the reference's.
"""
import ast
import importlib.abc
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))

_BINOPS = {ast.Add: "add", ast.Sub: "sub", ast.Mult: "mul", ast.Div: "div", ast.FloorDiv: "floordiv",
           ast.Mod: "mod", ast.Pow: "power", ast.BitXor: "bxor", ast.BitAnd: "band", ast.BitOr: "bor",
           ast.LShift: "shl", ast.RShift: "shr", ast.MatMult: "matmul"}
_BUILTINS = {"int": "ti_int", "float": "ti_float", "max": "ti_max", "min": "ti_min", "abs": "ti_abs",
             "pow": "ti_pow", "range": "ti_range"}
_PFX = "_tirt_"


def _rt(name):
    return ast.Name(id=_PFX + name, ctx=ast.Load())


def _call(name, *args):
    return ast.Call(func=_rt(name), args=list(args), keywords=[])


def _is_ti_attr(node, attr):
    return (isinstance(node, ast.Attribute) and node.attr == attr and isinstance(node.value, ast.Name)
            and node.value.id == "ti")


def _load(node):
    """copy of an assignment target usable as an expression"""
    n = ast.parse(ast.unparse(node), mode="eval").body
    return n


class _Body(ast.NodeTransformer):
    """rewrites ONE @ti.kernel / @ti.func body"""

    def __init__(self, argnames):
        self.scopes = [set(argnames)]

    # -- scoping ---------------------------------------------------------------------------
    def _defined(self, name):
        return any(name in s for s in self.scopes)

    def _block(self, stmts, names=()):
        self.scopes.append(set(names))
        out = []
        for s in stmts:
            r = self.visit(s)
            if isinstance(r, list):
                out.extend(r)
            elif r is not None:
                out.append(r)
        self.scopes.pop()
        return out

    # -- expressions -------------------------------------------------------------------------
    def visit_BinOp(self, node):
        l, r = self.visit(node.left), self.visit(node.right)
        return _call(_BINOPS[type(node.op)], l, r)

    def visit_Call(self, node):
        f = node.func
        if _is_ti_attr(f, "atomic_add") and isinstance(node.args[0], ast.Subscript):
            t = node.args[0]
            return _call("atomic_add_at", self.visit(t.value), self.visit(t.slice), self.visit(node.args[1]))
        if isinstance(f, ast.Name) and f.id in _BUILTINS and not self._defined(f.id):
            return ast.Call(func=_rt(_BUILTINS[f.id]), args=[self.visit(a) for a in node.args], keywords=[])
        return self.generic_visit(node)

    # -- statements ----------------------------------------------------------------------------
    def visit_Assign(self, node):
        value = self.visit(node.value)
        if len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if self._defined(name):
                value = _call("cast_like", ast.Name(id=name, ctx=ast.Load()), value)
            else:
                self.scopes[-1].add(name)
                value = _call("c", value)
            return ast.Assign(targets=[node.targets[0]], value=value)
        targets = []
        for t in node.targets:
            if isinstance(t, ast.Tuple):
                for el in t.elts:
                    if isinstance(el, ast.Name):
                        self.scopes[-1].add(el.id)
                targets.append(t)
            elif isinstance(t, ast.Name):
                self.scopes[-1].add(t.id)
                targets.append(t)
            else:
                targets.append(self.generic_visit(t))
        return ast.Assign(targets=targets, value=value)

    def visit_AugAssign(self, node):
        value = self.visit(node.value)
        op = _BINOPS[type(node.op)]
        t = node.target
        if isinstance(t, ast.Name):
            cur = ast.Name(id=t.id, ctx=ast.Load())
            return ast.Assign(targets=[ast.Name(id=t.id, ctx=ast.Store())],
                              value=_call("cast_like", cur, _call(op, ast.Name(id=t.id, ctx=ast.Load()), value)))
        store = self.generic_visit(t)
        load = self.visit(_load(t))
        return ast.Assign(targets=[store], value=_call(op, load, value))

    def visit_For(self, node):
        it = node.iter
        if isinstance(it, ast.Call) and isinstance(it.func, ast.Name) and it.func.id == "range":
            new_iter = self.visit(it)
        elif isinstance(it, ast.Call) and (_is_ti_attr(it.func, "ndrange") or _is_ti_attr(it.func, "static")
                                           or _is_ti_attr(it.func, "grouped")):
            new_iter = self.generic_visit(it)
        else:
            new_iter = _call("struct_iter", self.visit(it))
        names = [n.id for n in ast.walk(node.target) if isinstance(n, ast.Name)]
        body = self._block(node.body, names)
        return ast.For(target=node.target, iter=new_iter, body=body, orelse=[], type_comment=None)

    def visit_While(self, node):
        test = self.visit(node.test)
        return ast.While(test=test, body=self._block(node.body), orelse=[])

    def visit_If(self, node):
        test = self.visit(node.test)
        return ast.If(test=test, body=self._block(node.body),
                      orelse=self._block(node.orelse) if node.orelse else [])


class _Host(ast.NodeTransformer):
    """Python-scope code of the reference (module level, host helpers, constructors).  One thing is emulated there:
    the reference dates from NumPy 1.x (its .pyc files are cpython-37), where `np.float32 scalar <op> Python float` is
    evaluated in float64.  NumPy >= 2 (NEP 50) keeps float32, which changes e.g. HashGrid.py:47
    `int((max - min) / gridR + 1)` from 9 to 10 cells for a 0.45-wide scene.  Binary operators are routed through
    `host_op`, which promotes that one operand pairing like NumPy 1.x did and is the plain operator otherwise."""

    def visit_FunctionDef(self, node):
        if any(_is_ti_attr(d, "kernel") or _is_ti_attr(d, "func") for d in node.decorator_list):
            return node                      # already rewritten by _Module
        return self.generic_visit(node)

    def visit_BinOp(self, node):
        if type(node.op) not in _BINOPS or isinstance(node.op, ast.MatMult):
            return self.generic_visit(node)
        return _call("host_op", ast.Constant(_BINOPS[type(node.op)]), self.visit(node.left), self.visit(node.right))

    def visit_AugAssign(self, node):
        if type(node.op) not in _BINOPS or isinstance(node.op, ast.MatMult) or not isinstance(node.target, (ast.Name, ast.Subscript, ast.Attribute)):
            return self.generic_visit(node)
        value = self.visit(node.value)
        store = self.generic_visit(node.target) if not isinstance(node.target, ast.Name) else node.target
        load = self.visit(_load(node.target))
        return ast.Assign(targets=[store], value=_call("host_op", ast.Constant(_BINOPS[type(node.op)]), load, value))


class _Module(ast.NodeTransformer):
    def __init__(self, consts):
        self.consts = consts or {}
        self.depth = 0
        self.rewritten = []

    def visit_Module(self, node):
        for k, st in enumerate(node.body):
            if (isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name)
                    and st.targets[0].id in self.consts):
                st.value = ast.Constant(self.consts[st.targets[0].id])
        self.generic_visit(node)
        return node

    def visit_FunctionDef(self, node):
        is_ti = any(_is_ti_attr(d, "kernel") or _is_ti_attr(d, "func") for d in node.decorator_list)
        if not is_ti:
            return self.generic_visit(node)
        args = [a.arg for a in node.args.args]
        b = _Body(args)
        node.body = b._block(node.body)
        self.rewritten.append(node.name)
        return node


def transform_source(src, filename, consts=None):
    tree = ast.parse(src, filename)
    m = _Module(consts)
    tree = m.visit(tree)
    tree = _Host().visit(tree)
    ast.fix_missing_locations(tree)
    return compile(tree, filename, "exec"), m.rewritten


class RefLoader(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """serves top-level modules / the `kernels` package of `ref_dir`; `consts` = {module: {name: literal}}"""

    def __init__(self, ref_dir, consts=None):
        self.ref_dir = ref_dir
        self.consts = consts or {}
        self.loaded = []

    def _path(self, fullname):
        base = os.path.join(self.ref_dir, *fullname.split("."))
        if os.path.isfile(base + ".py"):
            return base + ".py", False
        if os.path.isdir(base) and "." not in fullname and fullname in ("kernels",):
            return base, True
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        p, is_pkg = self._path(fullname)
        if p is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, self, origin=p, is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [p]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import taichi as shim
        p, is_pkg = self._path(module.__name__)
        if is_pkg:
            return
        with open(p, "r") as f:
            src = f.read()
        code, rewritten = transform_source(src, p, self.consts.get(module.__name__))
        g = module.__dict__
        g["__file__"] = p
        for name in ("add", "sub", "mul", "div", "floordiv", "mod", "power", "bxor", "band", "bor", "shl", "shr",
                     "matmul", "c", "cast_like", "atomic_add_at", "struct_iter", "ti_int", "ti_float", "ti_max",
                     "ti_min", "ti_abs", "ti_pow", "ti_range", "host_op"):
            g[_PFX + name] = getattr(shim, name)
        self.loaded.append((module.__name__, rewritten))
        exec(code, g)

    # -- install / remove ---------------------------------------------------------------------
    def install(self):
        shim_dir = _HERE
        if shim_dir not in sys.path:
            sys.path.insert(0, shim_dir)
        sys.meta_path.insert(0, self)
        return self

    def remove(self):
        if self in sys.meta_path:
            sys.meta_path.remove(self)
        for name in [n for n, _ in self.loaded] + ["kernels"]:
            sys.modules.pop(name, None)
