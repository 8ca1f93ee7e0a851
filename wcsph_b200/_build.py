"""Builds libwcsph_b200.so in-tree with nvcc for sm_100a (no torch headers: plain C ABI)."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwcsph_b200.so")
SOURCES = ["api.cu", "grid.cu", "mgpu.cu", "sesph.cu", "dfsph.cu", "iisph.cu", "pcisph.cu", "canvas.cu", "mc.cu", "aniso.cu", "boundry.cu"]
# the Poisson-disk acceptance test compares a distance with particleRadius: IEEE divide / sqrt / asinf (no fast-math) keep its decisions
NO_FAST_MATH = {"boundry.cu"}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-use_fast_math",
              "-Xcompiler", "-fPIC", "--extended-lambda"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "wcsph_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: A/B variants for tools/ (e.g. defines=["WCSPH_UNROLL=2"], out="/path/lib_u2.so")"""
    lib = out or LIB
    if not out and not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" + ("_" + os.path.basename(lib).replace(".so", "") if out else ""))
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        flags = [f for f in NVCC_FLAGS if not (src in NO_FAST_MATH and f == "-use_fast_math")]
        cmd = [nvcc] + flags + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
