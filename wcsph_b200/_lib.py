"""ctypes binding of libwcsph_b200.so (include/wcsph_b200.h).  No CPU fallback: a missing
library or a missing CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WCSPH_LIB") or os.path.join(HERE, "libwcsph_b200.so")   # WCSPH_LIB: A/B builds (tools/)

SESPH, PCISPH, IISPH, DFSPH = 0, 1, 2, 3
SOLVER_ID = {"sesph": SESPH, "pcisph": PCISPH, "iisph": IISPH, "dfsph": DFSPH}
ABI_VERSION = 2

FLAG_BUCKET_OVERFLOW, FLAG_NEIGHBOR_OVERFLOW, FLAG_LIST_OVERFLOW, FLAG_ALIAS_OVERFLOW, FLAG_NAN = 1, 2, 4, 8, 16
FLAG_MC_OVERFLOW = 32
FLAG_MIGRATE_FAR = 64
FLAG_COMM_TIMEOUT = 128
FLAGS_FATAL = FLAG_BUCKET_OVERFLOW | FLAG_LIST_OVERFLOW | FLAG_ALIAS_OVERFLOW | FLAG_MIGRATE_FAR | FLAG_COMM_TIMEOUT


class Params(C.Structure):
    """struct wcsph_params"""
    _fields_ = [("searchR", C.c_float), ("m_k", C.c_float), ("m_l", C.c_float), ("m_k_raw", C.c_float),
                ("h3inv", C.c_float), ("kernel_style", C.c_int), ("coh_m_k", C.c_float), ("coh_m_c", C.c_float),
                ("adh_m_k", C.c_float), ("rho_L0", C.c_float), ("rho_S0", C.c_float), ("VL0", C.c_float),
                ("VS0", C.c_float), ("liqiudMass", C.c_float), ("gravity", C.c_float * 3),
                ("dim_coff", C.c_float), ("viscosity", C.c_float), ("viscosity_b", C.c_float),
                ("viscosity_err", C.c_float), ("tension_coff", C.c_float), ("tension_coff_b", C.c_float),
                ("viscosity_omega", C.c_float), ("vorticity_coff", C.c_float), ("vorticity_init", C.c_float),
                ("stiffness", C.c_float), ("pci_coff", C.c_float), ("omega_relax", C.c_float), ("eps", C.c_float),
                ("particleRadius", C.c_float), ("user_max_t", C.c_float), ("user_min_t", C.c_float)]


class Desc(C.Structure):
    """struct wcsph_desc"""
    _fields_ = [("abi_version", C.c_int), ("solver", C.c_int), ("count", C.c_int), ("liquid_count", C.c_int),
                ("hash_gridR", C.c_double), ("max_in_grid", C.c_int), ("max_neighbour", C.c_int),
                ("list_cap_liquid", C.c_int), ("list_cap_solid", C.c_int), ("cull_scale", C.c_float),
                ("min_boundary", C.c_float * 3), ("max_boundary", C.c_float * 3),
                ("world_size", C.c_int), ("rank", C.c_int), ("z_lo", C.c_int), ("z_hi", C.c_int),
                ("cap_own", C.c_int), ("cap_ghost", C.c_int), ("params", Params)]


# every entry point include/wcsph_b200.h declares: name -> (restype, argtypes)
_P, _I, _S = C.c_void_p, C.c_int, C.c_char_p
_CTX_ONLY = [
    "wcsph_hashgrid_update_grid",
    "wcsph_sesph_reset_param", "wcsph_sesph_update_advection_density", "wcsph_sesph_update_pressure",
    "wcsph_sesph_compute_force", "wcsph_sesph_integrator_sesph",
    "wcsph_dfsph_reset_param", "wcsph_dfsph_compute_density", "wcsph_dfsph_compute_dfsph_coff",
    "wcsph_dfsph_warmstart_divergence_vel", "wcsph_dfsph_begin_divergence_iter", "wcsph_dfsph_divergence_iter",
    "wcsph_dfsph_end_divergence_iter", "wcsph_dfsph_clear_nonpressure", "wcsph_dfsph_compute_tension",
    "wcsph_dfsph_init_viscosity_para", "wcsph_dfsph_compute_viscosity_force", "wcsph_dfsph_end_viscosity",
    "wcsph_dfsph_compute_vorticity", "wcsph_dfsph_cfl_max", "wcsph_dfsph_update_vel",
    "wcsph_dfsph_warmstart_pressure", "wcsph_dfsph_begin_pressure_iter", "wcsph_dfsph_pressure_iter",
    "wcsph_dfsph_end_pressure_iter", "wcsph_dfsph_update_pos",
    "wcsph_iisph_reset_param", "wcsph_iisph_compute_density", "wcsph_iisph_init_viscosity_para",
    "wcsph_iisph_compute_viscosity_force", "wcsph_iisph_combine_nonpressure", "wcsph_iisph_compute_advection",
    "wcsph_iisph_update_iter_info", "wcsph_iisph_update_pressure_force", "wcsph_iisph_update_pos",
    "wcsph_pcisph_reset_param", "wcsph_pcisph_compute_nonpressure_force", "wcsph_pcisph_compute_tension", "wcsph_pcisph_init_iter_info",
    "wcsph_pcisph_update_iter_info", "wcsph_pcisph_predict_density", "wcsph_pcisph_update_pos",
    "wcsph_sync", "wcsph_check",
]
_STEP = ["wcsph_sesph_step", "wcsph_dfsph_step", "wcsph_iisph_step", "wcsph_pcisph_step"]
SIGNATURES = {
    "wcsph_last_error": (C.c_char_p, []),
    "wcsph_abi_version": (_I, []),
    "wcsph_arena_bytes": (C.c_size_t, [C.POINTER(Desc)]),
    "wcsph_create": (_I, [C.POINTER(Desc), _P, C.c_size_t, _P, C.POINTER(_P)]),
    "wcsph_destroy": (None, [_P]),
    "wcsph_set_stream": (_I, [_P, _P]),
    "wcsph_set_params": (_I, [_P, C.POINTER(Params)]),
    "wcsph_upload_pos": (_I, [_P, _P]),
    "wcsph_block_size": (_I, [_P, C.POINTER(_I * 3)]),
    "wcsph_field_info": (_I, [_P, _S, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "wcsph_field_get": (_I, [_P, _S, _P, C.c_size_t]),
    "wcsph_field_set": (_I, [_P, _S, _P, C.c_size_t]),
    "wcsph_field_get_async": (_I, [_P, _S, _P, C.c_size_t]),
    "wcsph_field_set_async": (_I, [_P, _S, _P, C.c_size_t]),
    "wcsph_field_device": (_I, [_P, _S, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I)]),
    "wcsph_sorted_id_device": (_I, [_P, C.POINTER(_P)]),
    "wcsph_scalar_get": (_I, [_P, _S, C.POINTER(C.c_float)]),
    "wcsph_scalar_set": (_I, [_P, _S, C.c_float]),
    "wcsph_status": (_I, [_P, C.POINTER(C.c_uint32)]),
    "wcsph_iters": (_I, [_P, C.POINTER(_I * 3)]),
    "wcsph_launch_count": (C.c_longlong, [_P, _I]),
    "wcsph_iters_log": (_I, [_P, _P, _I, C.POINTER(_I)]),
    "wcsph_set_option": (_I, [_P, _S, _I]),
    "wcsph_set_iters": (_I, [_P, _I, _I, _I]),
    "wcsph_comm_unique_id": (_I, [_P, _S]),
    "wcsph_comm_init": (_I, [_P, _P, _S]),
    "wcsph_comm_info": (_I, [_P, _P]),
    "wcsph_comm_mailbox_handle": (_I, [_P, _P]),
    "wcsph_comm_mailbox_open": (_I, [_P, _P]),
    "wcsph_owned_count": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "wcsph_migration_counts": (_I, [_P, C.POINTER(C.c_longlong * 5)]),
    "wcsph_profile": (_I, [_P, _I]),
    "wcsph_profile_report": (_I, [_P, C.c_char_p, C.c_size_t]),
    "wcsph_hashgrid_neighbors_of": (_I, [_P, _I, _P, _I, C.POINTER(_I)]),
    "wcsph_pair_counts": (_I, [_P, C.POINTER(C.c_longlong * 4)]),
}
for _n in _CTX_ONLY:
    SIGNATURES[_n] = (_I, [_P])
for _n in _STEP:
    SIGNATURES[_n] = (_I, [_P, _I])



class McGrid(C.Structure):
    """struct wcsph_mc_grid."""
    _fields_ = [("gridR", C.c_double), ("isolevel", C.c_float), ("max_in_grid", C.c_int), ("liqiudMass", C.c_float),
                ("min_boundary", C.c_float * 3), ("block", C.c_int * 3)]


SIGNATURES["wcsph_mc_workspace_bytes"] = (C.c_size_t, [C.POINTER(McGrid), _I])
SIGNATURES["wcsph_mc_update_grid"] = (_I, [_P, C.POINTER(McGrid), _P, C.c_size_t])
SIGNATURES["wcsph_mc_cal_surface_point"] = (_I, [_P, C.POINTER(McGrid), _P, C.c_size_t, _P])
SIGNATURES["wcsph_mc_marching_cube"] = (_I, [_P, C.POINTER(McGrid), _P, C.c_size_t, _P, _P, _I, C.POINTER(_I)])
SIGNATURES["wcsph_pd_aniso_workspace_bytes"] = (C.c_size_t, [_P])
SIGNATURES["wcsph_pd_compute_color_map"] = (_I, [_P, _P, C.c_size_t, _P, _P])
SIGNATURES["wcsph_pd_cal_anistropic_kernel"] = (_I, [_P, C.c_float, _P, C.c_size_t, _P, _P])
SIGNATURES["wcsph_mc_cal_surface_point_anistropic"] = (_I, [_P, C.POINTER(McGrid), _P, C.c_size_t, _P, _P, _P])
SIGNATURES["wcsph_canvas_clear"] = (_I, [_P, _P, _I, _I])
SIGNATURES["wcsph_canvas_draw_particle"] = (_I, [_P, _P, _P, _I, _I, _I, _P])
SIGNATURES["wcsph_canvas_resolve"] = (_I, [_P, _P, _I, _I, _P, _P])



class BdDesc(C.Structure):
    """struct wcsph_bd_desc."""
    _fields_ = [("n", C.c_int), ("padding", C.c_int), ("hash_size", C.c_int), ("phase_vec_max", C.c_int), ("sample_cap", C.c_int),
                ("radius", C.c_float), ("gridR", C.c_float), ("min_point", C.c_float * 3)]


SIGNATURES["wcsph_bd_workspace_bytes"] = (C.c_size_t, [C.POINTER(BdDesc)])
SIGNATURES["wcsph_bd_init_point_set"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _P, _P, _I, C.c_float, C.c_uint, _P])
SIGNATURES["wcsph_bd_set_points"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _P, _P, _P])
SIGNATURES["wcsph_bd_bitonic_sort"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _P])
SIGNATURES["wcsph_bd_build_hmap"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _P])
SIGNATURES["wcsph_bd_sample"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _P, _I, _I, _P])
SIGNATURES["wcsph_bd_get"] = (_I, [C.POINTER(BdDesc), _P, C.c_size_t, _S, _P, C.c_size_t, _P])

_lib = None


class WcsphError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and bind every declared symbol; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WcsphError("libwcsph_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                         "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)          # AttributeError if the symbol is missing
        f.restype = res
        f.argtypes = args
    if L.wcsph_abi_version() != ABI_VERSION:
        raise WcsphError("ABI mismatch: library %d, binding %d" % (L.wcsph_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise WcsphError("libwcsph_b200: %s (rc=%d)" % (load().wcsph_last_error().decode(), rc))
