"""ctypes loader for liblpmx.so (the C ABI declared in include/lpmx.h).

The product path FAILS LOUDLY when the CUDA library is missing or no B200 is present: there is no
CPU fallback anywhere in this package (the mesh generator is host code in the reference too and is
the only part usable without a GPU).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblpmx.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ubyte_p = ctypes.POINTER(ctypes.c_ubyte)
c_long_p = ctypes.POINTER(ctypes.c_long)
vp = ctypes.c_void_p

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NO_DEVICE, ERR_COMM, ERR_UNSUPPORTED, ERR_STATE = -1, -2, -3, -4, -5, -6, -7
LAYOUT_RIGHT, LAYOUT_LEFT = 0, 1
SEED_ICOS_TRI_SPHERE, SEED_CUBED_SPHERE, SEED_QUAD_RECT, SEED_TRI_HEX = 0, 1, 2, 3


class SwePassive(ctypes.Structure):
    """lpmx_swe_passive_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("xyz", "vort", "div", "depth", "surf", "bottom", "vel", "ddot", "laps")]


class SweActive(ctypes.Structure):
    """lpmx_swe_active_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("xyz", "vort", "div", "area", "mass", "depth", "surf", "bottom", "vel",
                                               "ddot", "laps", "mask")]


PLANE_PASSIVE_FIELDS = ("xy", "vort", "div", "depth", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2", "du2dx1",
                        "du2dx2", "laps", "psi", "phi")
PLANE_ACTIVE_FIELDS = ("xy", "vort", "div", "area", "mass", "depth", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2",
                       "du2dx1", "du2dx2", "laps", "psi", "phi")
PLANE_SUM_FIELDS = ("vel", "ddot", "du1dx1", "du1dx2", "du2dx1", "du2dx2", "laps", "psi", "phi")


class PlaneSwePassive(ctypes.Structure):
    """lpmx_plane_swe_passive_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in PLANE_PASSIVE_FIELDS]


class PlaneSweActive(ctypes.Structure):
    """lpmx_plane_swe_active_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in PLANE_ACTIVE_FIELDS + ("mask",)]


class PlaneSweSums(ctypes.Structure):
    """lpmx_plane_swe_sums_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in PLANE_SUM_FIELDS]


# lpmx_swe_laplacian_fn
SWE_LAPLACIAN_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long)


class GmlsParams(ctypes.Structure):
    """lpmx_gmls_params_t == gmls::Params (src/lpm_compadre.hpp:23-60)"""
    _fields_ = [("eps_multiplier", ctypes.c_double), ("samples_order", ctypes.c_int), ("manifold_order", ctypes.c_int),
                ("samples_weight_pwr", ctypes.c_double), ("manifold_weight_pwr", ctypes.c_double),
                ("ambient_dim", ctypes.c_int), ("topo_dim", ctypes.c_int), ("min_neighbors", ctypes.c_int)]


class FlagDesc(ctypes.Structure):
    """lpmx_flag_desc_t: one refinement-flag functor of src/mesh/lpm_refinement_flags.hpp"""
    _fields_ = [("kind", ctypes.c_int), ("n_faces", ctypes.c_int), ("n_verts", ctypes.c_int), ("n_face_verts", ctypes.c_int),
                ("face_vals", ctypes.c_void_p), ("area", ctypes.c_void_p), ("vert_vals", ctypes.c_void_p),
                ("face_verts", ctypes.c_void_p), ("vert_lag", ctypes.c_void_p), ("ndim", ctypes.c_int),
                ("layout", ctypes.c_int), ("ld", ctypes.c_long), ("mask", ctypes.c_void_p), ("tol", ctypes.c_double)]


class GmlsProvider(ctypes.Structure):
    """lpmx_gmls_provider_t"""
    _fields_ = [("handle", ctypes.c_void_p), ("params", GmlsParams)]


class LpmxError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {error_name(code)} ({code}) {detail}")


_lib = None


def lib():
    """Load liblpmx.so; raises if it has not been built (python -m lpm_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m lpm_b200.build` "
                "(lpm_b200 has no CPU or pure-Python fallback)")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _declare(L)
        _lib = L
    return _lib


def error_name(code):
    return lib().lpmx_error_name(code).decode()


def _declare(L):
    d, i, l = ctypes.c_double, ctypes.c_int, ctypes.c_long
    L.lpmx_version_string.restype = ctypes.c_char_p
    L.lpmx_error_name.restype = ctypes.c_char_p
    L.lpmx_error_name.argtypes = [i]
    L.lpmx_last_error_string.restype = ctypes.c_char_p
    L.lpmx_last_error_string.argtypes = [vp]
    sig = {
        "lpmx_mesh_max_allocations": [i, i, c_int_p, c_int_p, c_int_p],
        "lpmx_mesh_create": [i, i, d, ctypes.POINTER(vp)],
        "lpmx_mesh_destroy": [vp],
        "lpmx_mesh_sizes": [vp, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p],
        "lpmx_mesh_array": [vp, i, ctypes.POINTER(vp), c_long_p, c_int_p],
        "lpmx_mesh_update_array": [vp, i, vp, l],
        "lpmx_mesh_divide_flagged_faces": [vp, vp, i, i, i, c_int_p, c_int_p],
        "lpmx_mesh_leaf_edges_from_parent": [vp, i, vp, i, c_int_p],
        "lpmx_mesh_ccw_edges_around_face": [vp, i, vp, i, c_int_p],
        "lpmx_mesh_ccw_adjacent_faces": [vp, i, vp, i, c_int_p],
        "lpmx_mesh_neighbors_flag": [vp, vp, i, i, c_int_p],
        "lpmx_mesh_locate": [vp, i, vp, i, vp, vp],
        "lpmx_refine_flag_max": [vp, ctypes.POINTER(FlagDesc), ctypes.POINTER(ctypes.c_double)],
        "lpmx_refine_flag": [vp, ctypes.POINTER(FlagDesc), i, i, i, vp, c_int_p],
        "lpmx_create": [ctypes.POINTER(vp), i],
        "lpmx_destroy": [vp],
        "lpmx_sync": [vp],
        "lpmx_stream": [vp, ctypes.POINTER(vp)],
        "lpmx_launch_count": [vp, c_long_p],
        "lpmx_copy": [vp, vp, vp, l],
        "lpmx_profile_enable": [vp, i],
        "lpmx_profile_read": [vp, c_long_p, c_double_p, c_double_p],
        "lpmx_set_partition": [vp, i, i],
        "lpmx_comm_unique_id": [vp],
        "lpmx_comm_init": [vp, vp, i, i],
        "lpmx_comm_enable_peer_exchange": [vp, i],
        "lpmx_comm_peer_exchange_enabled": [vp, c_int_p, c_int_p],
        "lpmx_pair_sum_const_stream": [vp, i],
        "lpmx_const_stream_split": [i, i, i, c_int_p, c_int_p, c_int_p, c_int_p, c_double_p, c_double_p],
        "lpmx_const_stream_launch_count": [vp, c_long_p],
        "lpmx_fp64_peak_tflops": [vp, c_double_p, c_double_p],
        "lpmx_bve_velocity": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, i, i, vp],
        "lpmx_bve_streamfn": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, i, i, vp],
        "lpmx_set_io_sharded": [vp, i],
        "lpmx_local_targets": [vp, i, i, vp, vp, vp, vp],
        "lpmx_bve_solve": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, i, i, vp, vp],
        "lpmx_ic2d_sums": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, i, d, i, vp, vp],
        "lpmx_swe_sphere_sums": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, vp, i, d, i, i, vp, vp, vp],
        "lpmx_bve_rk4_step": [vp, d, d, i, vp, vp, vp, i, vp, vp, vp, vp, vp, i, l, l, i],
        "lpmx_bve_solver_create": [vp, i, i, ctypes.POINTER(vp)],
        "lpmx_bve_solver_destroy": [vp],
        "lpmx_bve_solver_set_state": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, l, l],
        "lpmx_bve_solver_get_state": [vp, vp, vp, vp, vp, vp, vp, i, l, l],
        "lpmx_bve_solver_init_velocity": [vp],
        "lpmx_bve_solver_stream_fn": [vp, vp, vp],
        "lpmx_bve_solver_advance": [vp, d, d, i],
        "lpmx_bve_solver_interactions_per_eval": [vp, c_double_p, c_double_p],
        "lpmx_ic2d_rk2_step": [vp, d, d, d, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, l, l, i],
        "lpmx_ic2d_solver_create": [vp, i, i, d, ctypes.POINTER(vp)],
        "lpmx_ic2d_solver_destroy": [vp],
        "lpmx_ic2d_solver_set_state": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, l, l],
        "lpmx_ic2d_solver_get_state": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, l, l],
        "lpmx_ic2d_solver_init_direct_sums": [vp],
        "lpmx_ic2d_solver_advance": [vp, d, d, i],
        "lpmx_ic2d_solver_lazy_stream_fn": [vp, i],
        "lpmx_ic2d_totals": [vp, i, vp, vp, i, l, vp, vp, c_double_p, c_double_p, c_double_p],
        "lpmx_ic2d_solver_totals": [vp, c_double_p, c_double_p, c_double_p],
        "lpmx_err_norms": [vp, i, i, vp, vp, i, l, vp, c_double_p, c_double_p, c_double_p],
        "lpmx_gather_mesh_data": [vp, i, i, i, vp, l, i, vp, l, vp, vp, l, c_int_p],
        "lpmx_scatter_mesh_data": [vp, i, i, vp, l, i, vp, l, i, vp, l, vp],
        "lpmx_gmls_params_init": [ctypes.POINTER(GmlsParams), i],
        "lpmx_gmls_sphere_laplacian": [vp, ctypes.POINTER(GmlsParams), i, vp, i, l, vp, vp, vp, vp],
        "lpmx_gmls_sphere_interpolate": [vp, ctypes.POINTER(GmlsParams), i, vp, i, l, i, vp, i, vp, i, l, vp],
        "lpmx_gmls_swe_laplacian": [vp, i, vp, i, vp, vp, vp, i, vp, vp, vp, vp, l],
        "lpmx_ftle": [vp, i, i, vp, vp, i, l, i, vp, vp, l, vp, i, vp, vp, c_double_p],
        "lpmx_swe_rk2_step": [vp, d, d, d, d, i, ctypes.POINTER(SwePassive), i, ctypes.POINTER(SweActive), i, l, l,
                              SWE_LAPLACIAN_FN, vp, i],
        "lpmx_swe_solver_create": [vp, i, i, d, ctypes.POINTER(vp)],
        "lpmx_swe_solver_destroy": [vp],
        "lpmx_swe_solver_set_state": [vp, ctypes.POINTER(SwePassive), ctypes.POINTER(SweActive), i, l, l],
        "lpmx_swe_solver_get_state": [vp, ctypes.POINTER(SwePassive), ctypes.POINTER(SweActive), i, l, l],
        "lpmx_swe_solver_set_laplacian": [vp, vp, vp],
        "lpmx_swe_solver_init_direct_sums": [vp, i],
        "lpmx_swe_solver_advance": [vp, d, d, d, SWE_LAPLACIAN_FN, vp, i],
        "lpmx_ic2d_plane_sums": [vp, vp, i, l, i, vp, i, l, vp, vp, vp, i, d, i, vp, vp],
        "lpmx_ic2d_plane_rk2_step": [vp, d, d, d, d, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, l, l, i],
        "lpmx_swe_plane_sums": [vp, vp, i, l, vp, i, vp, i, l, vp, vp, vp, vp, vp, i, d, d, i, i,
                                ctypes.POINTER(PlaneSweSums)],
        "lpmx_swe_plane_rk4_step": [vp, d, d, d, d, d, d, i, i, ctypes.POINTER(PlaneSwePassive), i,
                                    ctypes.POINTER(PlaneSweActive), i, l, l, i],
        "lpmx_plane_swe_solver_create": [vp, i, i, d, d, i, ctypes.POINTER(vp)],
        "lpmx_plane_swe_solver_destroy": [vp],
        "lpmx_plane_swe_solver_set_state": [vp, ctypes.POINTER(PlaneSwePassive), ctypes.POINTER(PlaneSweActive), i, l, l],
        "lpmx_plane_swe_solver_get_state": [vp, ctypes.POINTER(PlaneSwePassive), ctypes.POINTER(PlaneSweActive), i, l, l],
        "lpmx_plane_swe_solver_init_direct_sums": [vp, i],
        "lpmx_plane_swe_solver_advance": [vp, d, d, d, d, i],
        "lpmx_plane_swe_solver_interactions_per_eval": [vp, c_double_p, c_double_p],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i


def declared_symbols():
    """Every function include/lpmx.h declares (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "lpmx.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lpmx_[a-z0-9_]+)\s*\(", text)))
