"""boundry -- the reference's boundary pre-processing tool (boundry.py) on the GPU.

Parallel Poisson-disk sampling of a triangle mesh (Bowers et al. 2010): `<name>.obj` in, `<name>_boundry.obj` out -- the solid
point cloud the solver scripts load (`init_particle("model/box_boundry.obj")`, dfsph.py:597).  Same module surface as the
reference script: constants (`particleRadius`, `gridR`, `phase_block_size`, `hash_sample_size`, `trial_total`), `get_pot_num`,
`loadObj(filename)`, `init_point_set()`, `gpu_bitonic_sort()`, `build_hmap()`, `detect_hmap()`, `possion_disk_sample(pg, trial,
pg_count)`, and `main(inputfile)` = the script body (boundry.py:409-457) without the window.  Every former @ti.kernel is one call
into libwcsph_b200 (csrc/boundry.cu).  Nothing runs at import.

`init_point_set()` draws from the library's own counter-based generator (ti.random's stream is not reproducible outside Taichi);
`set_initial_points(pos, face_id)` injects a point set instead -- from there on the result is the reference's, point for point
(tests/test_boundry_gpu.py against the executed reference).  No CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from . import _lib

particleRadius = 0.025                          # boundry.py:21-22
gridR = particleRadius / math.sqrt(3.0)
pi = 3.1415926                                  # :54
phase_block_size = 27                           # :59
hash_sample_size = 5                            # :62
trial_total = 10                                # :417

faceNum = 0
numInitialPoints = 0
padding_num = 0
maxArea = 0.0
totalArea = 0.0
phase_vec_max = 0
hash_map_size = 0
min_point = [100000.0, 100000.0, 100000.0]
max_point = [-100000.0, -100000.0, -100000.0]

_s = {}        # device state of the loaded mesh


def get_pot_num(num):
    """boundry.py:82-86."""
    m = 1
    while m < num:
        m = m << 1
    return m >> 1


def loadObj(filename):
    """boundry.py:99-193: triangle soup, per-vertex copies of the face normals, areas; sizes of every table."""
    global faceNum, numInitialPoints, padding_num, maxArea, totalArea, phase_vec_max, hash_map_size, min_point, max_point
    import torch
    if not torch.cuda.is_available():
        raise _lib.WcsphError("wcsph_b200.boundry needs a CUDA device; there is no CPU fallback")
    vertices, faces = [], []
    for line in open(filename, "r"):
        if line.startswith('#'):
            continue
        values = line.split()
        if not values:
            continue
        if values[0] == 'v':
            vertices.append(list(map(float, values[1:4])))
        elif values[0] == 'f':
            for v in values[1:]:
                faces.append(int(v.split('/')[0]))
    faceNum = len(faces) // 3
    V = np.asarray(vertices, dtype=np.float64)
    F = np.asarray(faces[:3 * faceNum], dtype=np.int64).reshape(faceNum, 3) - 1
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    used = np.concatenate([a, b, c])
    min_point = [min(100000.0, float(used[:, k].min())) for k in range(3)]
    max_point = [max(-100000.0, float(used[:, k].max())) for k in range(3)]
    arrV = np.stack([a, b, c], axis=1).astype(np.float32)                      # [face][3][3]
    n = np.cross(b - a, c - a)
    ln = np.linalg.norm(n, axis=1)
    arrArea = (ln * 0.5).astype(np.float32)                                     # :140
    nn = (n / ln[:, None]).astype(np.float32)
    arrN = np.repeat(nn, 3, axis=0)                                             # one normal per VERTEX (:143-145)
    totalArea, maxArea = 0.0, 0.0
    for i in range(faceNum):                                                    # float32 areas accumulated in Python floats (:147-148)
        totalArea += float(arrArea[i])
        maxArea = max(float(arrArea[i]), maxArea)
    circleArea = pi * particleRadius * particleRadius
    numInitialPoints = int(40.0 * (totalArea / circleArea))                     # :151
    padding_num = get_pot_num(numInitialPoints) << 1
    phase_vec_max = numInitialPoints // 8
    hash_map_size = numInitialPoints * 3
    d = _lib.BdDesc()
    d.n, d.padding, d.hash_size, d.phase_vec_max, d.sample_cap = numInitialPoints, padding_num, hash_map_size, phase_vec_max, hash_sample_size
    d.radius, d.gridR = particleRadius, gridR
    for k in range(3):
        d.min_point[k] = min_point[k]
    L = _lib.load()
    nbytes = L.wcsph_bd_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise _lib.WcsphError("boundry: mesh too small / bad sizes (n = %d)" % numInitialPoints)
    _s.clear()
    _s.update(desc=d, work=torch.zeros(nbytes, dtype=torch.uint8, device="cuda"),
              tri_v=torch.from_numpy(arrV.reshape(-1)).cuda(), tri_n=torch.from_numpy(arrN.reshape(-1)).cuda(),
              tri_a=torch.from_numpy(arrArea).cuda(), stream=torch.cuda.current_stream(), sample_count=0,
              tri_vertices=arrV.reshape(-1, 3), tri_normal=arrN, tri_area=arrArea)
    torch.cuda.synchronize()
    return numInitialPoints


def _args():
    if not _s:
        raise _lib.WcsphError("boundry: loadObj() first")
    return C.byref(_s["desc"]), C.c_void_p(_s["work"].data_ptr()), _s["work"].numel()


def _stream():
    return C.c_void_p(_s["stream"].cuda_stream)


def init_point_set(seed=1):
    """boundry.py:223-247 with the library's generator."""
    d, w, n = _args()
    _lib.check(_lib.load().wcsph_bd_init_point_set(d, w, n, C.c_void_p(_s["tri_v"].data_ptr()), C.c_void_p(_s["tri_a"].data_ptr()),
                                                   faceNum, C.c_float(maxArea), C.c_uint(seed), _stream()))


def set_initial_points(init_pos, init_id):
    """inject the initial point set (what init_point_set would have drawn): positions [n,3] f32, face ids [n] i32."""
    d, w, n = _args()
    p = np.ascontiguousarray(init_pos, np.float32)
    i = np.ascontiguousarray(init_id, np.int32)
    if p.shape != (numInitialPoints, 3) or i.shape != (numInitialPoints,):
        raise ValueError("set_initial_points wants %d points" % numInitialPoints)
    _lib.check(_lib.load().wcsph_bd_set_points(d, w, n, p.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), _stream()))


def gpu_bitonic_sort():
    """boundry.py:208-219."""
    d, w, n = _args()
    _lib.check(_lib.load().wcsph_bd_bitonic_sort(d, w, n, _stream()))


def build_hmap():
    """boundry.py:250-271."""
    d, w, n = _args()
    _lib.check(_lib.load().wcsph_bd_build_hmap(d, w, n, _stream()))


def fetch(name):
    """device table -> numpy (see wcsph_bd_get)."""
    d, w, n = _args()
    D = _s["desc"]
    V = max(D.phase_vec_max, 1)
    shapes = {"cell": ((D.padding, 4), np.int32), "pos": ((D.padding, 4), np.float32), "start_index": ((D.hash_size,), np.int32),
              "hcell": ((D.hash_size, 4), np.int32), "hash_trace": ((D.n,), np.int32), "phase_group_count": ((27,), np.int32),
              "phase_group": ((27, V, 4), np.int32), "sample_count": ((D.hash_size,), np.int32), "sample": ((D.hash_size, D.sample_cap), np.int32),
              "possion_sample": ((D.n, 3), np.float32), "selected": ((D.n,), np.int32), "counters": ((4,), np.int32)}
    shp, dt = shapes[name]
    out = np.zeros(shp, dt)
    _lib.check(_lib.load().wcsph_bd_get(d, w, n, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes, _stream()))
    return out


def detect_hmap():
    """boundry.py:88-97: occupied hash slots counted two ways."""
    tr = fetch("hash_trace")
    cnt = fetch("counters")
    cpu = int(np.count_nonzero(tr))
    ok = cpu == int(cnt[1]) or cpu == int(cnt[1]) - 1          # a head that hashes to slot 0 leaves hash_trace at 0 (as in the reference)
    print("hash map ok!" if ok else "hash map error!", "cpu:", cpu, "gpu:", int(cnt[1]))
    return ok


def possion_disk_sample(pg, trial, pg_count=None):
    """boundry.py:390-407, one launch (pg_count is read on the device)."""
    d, w, n = _args()
    _lib.check(_lib.load().wcsph_bd_sample(d, w, n, C.c_void_p(_s["tri_n"].data_ptr()), int(pg), int(trial), _stream()))


def launch_order(trials=None, phases=None):
    """(phase, trial) in the order of the reference's main loop (boundry.py:421-457): phase_process is incremented BEFORE the first
    launch, so trial 0 never visits phase group 0."""
    trials = trial_total if trials is None else trials
    phases = phase_block_size if phases is None else phases
    out, phase, trial = [], 0, 0
    while True:
        if trial < trials:
            phase += 1
            if phase % phases == 0:
                trial += 1
                phase = 0
        if trial < trials:
            out.append((phase, trial))
        else:
            return out


def sample_all():
    for pg, trial in launch_order():
        possion_disk_sample(pg, trial)
    n = int(fetch("counters")[0])
    _s["sample_count"] = n
    return fetch("possion_sample")[:n]


def export_obj(path, pos):
    """boundry.py:445-451: `v x y z` per sample (Python's repr of the float32 values, like print does)."""
    with open(path, "w") as fo:
        for p in pos:
            fo.write("v %s %s %s\n" % (p[0], p[1], p[2]))
    return path


def main(inputfile="box", seed=1, init=None):
    """boundry.py:409-457: load, draw the initial points (or take `init` = (pos, face_id)), sort, hash, sample, write."""
    loadObj(inputfile + ".obj")
    if init is None:
        init_point_set(seed)
    else:
        set_initial_points(*init)
    gpu_bitonic_sort()
    build_hmap()
    detect_hmap()
    pos = sample_all()
    print("write obj")
    export_obj(inputfile + "_boundry.obj", pos)
    return pos
