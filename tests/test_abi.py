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


def test_host_kernel_mirrors_match_closed_forms():
    """wcsph_b200/kernels/*: the constants are the product surface; the evaluators follow the reference formulas"""
    import math
    import numpy as np
    from wcsph_b200.kernels.CubicKernel import CubicKernel
    from wcsph_b200.kernels.CohesionKernel import CohesionKernel
    from wcsph_b200.kernels.AdhesionKernel import AdhesionKernel
    h = 0.1
    k = CubicKernel(h)
    assert (k.m_k, k.m_l, k.h3) == (8.0 / math.pi, 48.0 / math.pi, 1.0 / (h * h * h))
    assert float(k.Cubic_W_norm(0.0)) == pytest.approx(8 / (math.pi * h ** 3), rel=1e-6)
    assert float(k.Cubic_W_norm(h / 2)) == pytest.approx(0.25 * 8 / (math.pi * h ** 3), rel=1e-6)
    assert float(k.Cubic_W_norm(1.01 * h)) == 0.0 and np.all(k.CubicGradW([0.0, 0.0, 0.0]) == 0.0)
    g = k.CubicGradW([0.03, 0.0, 0.0])
    q = 0.3
    assert float(g[0]) == pytest.approx(48 / (math.pi * h ** 3) * q * (3 * q - 2) / h, rel=1e-5)
    c = CohesionKernel(h)
    assert float(c.Cubic_W_norm(0.07)) == pytest.approx(32 / (math.pi * h ** 9) * 0.03 ** 3 * 0.07 ** 3, rel=1e-4)
    a = AdhesionKernel(h)
    assert float(a.Cubic_W_norm(0.075)) == pytest.approx(0.007 / h ** 3.25 * (-4 * 0.075 ** 2 / h + 6 * 0.075 - 2 * h) ** 0.25, rel=1e-4)
    assert float(a.Cubic_W_norm(0.04)) == 0.0


def test_bench_roofline_groups_template_instantiations():
    """bench.py: the dominant kernel is a kernel FAMILY; bytes and ncu traffic are launch-weighted over its instantiations."""
    import bench
    rows = {"(k_dfsph_drho<0, false, false, true>)": (10, 1.0), "(k_dfsph_drho<1, false, true, false>)": (5, 0.5),
            "k_build_lists": (5, 1.2), "k_finalize": (40, 0.3)}
    ab = {"(k_dfsph_drho<0, false, false, true>)": 100e6, "(k_dfsph_drho<1, false, true, false>)": 130e6, "k_build_lists": 20e6}
    ncu = {"k_dfsph_drho<0, 0, 0, 1>": 180e6, "k_build_lists": 130e6}
    r = bench.roofline_from_rows(rows, ab, ncu, 6500.0, "test", 5)
    assert r["kernel"] == "k_dfsph_drho" and r["launches_per_step"] == 3.0
    assert r["achieved"] == pytest.approx((10 * 100e6 + 5 * 130e6) / 1.5e-3 / 1e9)
    assert r["frac"] == pytest.approx(r["achieved"] / 6500.0) and r["traffic"] == 180e6
    assert r["share_of_step"] == pytest.approx(1.5 / 3.0)
    assert [x["kernel"] for x in r["next"]] == ["k_build_lists"]          # k_finalize has no algorithmic-byte entry


def test_anisotropic_branch_is_part_of_the_surface():
    """ParticleData.compute_color_map / cal_anistropic_kernel / export_kernel and MCGrid.cal_surface_point_anistropic exist with the
    reference's names (ParticleData.py:187-317, MarchingCubeGrid.py:215-246) and fail loudly without a device (no CPU fallback)."""
    from wcsph_b200.ParticleData import ParticleData
    from wcsph_b200.MarchingCubeGrid import MCGrid
    from wcsph_b200 import _lib
    pd = ParticleData(0.025)
    for name in ("compute_color_map", "cal_anistropic_kernel", "export_kernel"):
        assert callable(getattr(pd, name))
    assert callable(getattr(MCGrid, "cal_surface_point_anistropic"))
    for name in ("wcsph_pd_aniso_workspace_bytes", "wcsph_pd_compute_color_map", "wcsph_pd_cal_anistropic_kernel",
                 "wcsph_mc_cal_surface_point_anistropic", "wcsph_check", "wcsph_pair_counts", "wcsph_migration_counts"):
        assert name in _lib.SIGNATURES and hasattr(_lib.load(), name)


def test_header_is_plain_c_and_cpp():
    """the drop-in boundary is a C ABI: include/wcsph_b200.h must compile as C99 and as C++ on its own (no torch / CUDA types)"""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "wcsph_b200.h")
    for cc, lang, std in (("gcc", "c", "-std=c99"), ("g++", "c++", "-std=c++11")):
        if not shutil.which(cc):
            pytest.skip(cc + " not installed")
        r = subprocess.run([cc, std, "-Wall", "-fsyntax-only", "-x", lang, hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    import re
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)          # comments mention torch as an example caller
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "#include <cuda" not in code      # streams travel as void*
    assert set(re.findall(r"#include <([^>]+)>", code)) == {"stddef.h", "stdint.h"}
