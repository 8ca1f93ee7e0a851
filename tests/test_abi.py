"""CPU-side checks of the drop-in boundary: the C-ABI library is built in-tree, loads, and
exports every symbol include/wcsph_b200.h declares; the ctypes structs match the header.
No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from wcsph_b200 import _build
    return _build.build()


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "wcsph_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(wcsph_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(built)
    names = _header_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_binding_covers_header(built):
    from wcsph_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    L = _lib.load()
    assert L.wcsph_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    from wcsph_b200 import _lib
    # wcsph_params: 33 4-byte members; wcsph_desc: 4 ints, double, 4 ints, float, 2x float[3], params
    assert C.sizeof(_lib.Params) == 33 * 4
    assert _lib.Desc.hash_gridR.offset == 16 and _lib.Desc.params.offset == 16 + 8 + 16 + 4 + 24 + 24


def test_bad_descriptor_is_rejected_without_gpu(built):
    from wcsph_b200 import _lib
    L = _lib.load()
    d = _lib.Desc()
    d.abi_version = 999
    assert L.wcsph_arena_bytes(C.byref(d)) == 0
    assert b"abi_version" in L.wcsph_last_error()


def test_no_cpu_fallback():
    """the product path must fail loudly without a CUDA device"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wcsph_b200 import dfsph, _lib
    with pytest.raises(_lib.WcsphError):
        dfsph.init_particle("box_boundry")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wcsph_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "wcsph_oracle" not in src, f
