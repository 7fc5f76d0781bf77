"""numpy restatement of the reference's refinement-flag functors and Refinement::iterate.  TEST INFRASTRUCTURE ONLY.

  ScalarMaxFlag            /root/reference/src/mesh/lpm_refinement_flags.hpp:132-183
  ScalarIntegralFlag       :185-229
  ScalarVariationFlag      :231-310
  FlowMapVariationFlag     :55-130
  Refinement::iterate      /root/reference/src/mesh/lpm_refinement.hpp:28-41

Pinned against the reference header compiled in place (oracle/ref_flags_driver.cpp -> oracle/_ref/liblpm_ref.so) and against
tests/golden/ref_flags.npz, which tests/golden/make_ref_flags_golden.py generates from that build.

Quirks kept: the maxima of ScalarMaxFlag / ScalarIntegralFlag run over ALL faces (divided ones included; their area is 0 but
their field value is whatever the caller left there), those of the two variation flags over undivided faces only; every
comparison is strict (>); flags are only switched on; the maximum of an empty set is Kokkos' identity, -DBL_MAX.
"""
import numpy as np

LOWEST = -np.finfo(np.float64).max
KINDS = ("scalar_max", "scalar_integral", "scalar_variation", "flow_map_variation")


def flag_values(kind, face_mask, face_vals=None, area=None, vert_vals=None, face_verts=None, vert_lag=None):
    """The quantity the functor compares with its tolerance, per face (0 where the functor never looks)."""
    mask = np.asarray(face_mask) != 0
    n = mask.shape[0]
    if kind == "scalar_max":
        return np.abs(np.asarray(face_vals, dtype=np.float64))
    if kind == "scalar_integral":
        return np.abs(np.asarray(face_vals, dtype=np.float64)) * np.asarray(area, dtype=np.float64)
    out = np.zeros(n)
    live = ~mask
    fv = np.asarray(face_verts)[live]
    if kind == "scalar_variation":
        vals = np.concatenate([np.asarray(face_vals, dtype=np.float64)[live][:, None], np.asarray(vert_vals)[fv]], axis=1)
        out[live] = vals.max(axis=1) - vals.min(axis=1)
        return out
    if kind == "flow_map_variation":
        lag = np.asarray(vert_lag, dtype=np.float64)[fv]  # (n_live, nfv, ndim)
        ext = lag.max(axis=1) - lag.min(axis=1)
        dsum = np.zeros(ext.shape[0])
        for k in range(ext.shape[1]):  # sequential sum over the coordinates, as coded
            dsum = dsum + ext[:, k]
        out[live] = dsum
        return out
    raise ValueError(kind)


def flag_max(kind, face_mask, **arrays):
    """set_tol_from_relative_value(): the maximum the relative tolerance multiplies."""
    v = flag_values(kind, face_mask, **arrays)
    if kind in ("scalar_variation", "flow_map_variation"):
        v = v[np.asarray(face_mask) == 0]
    return float(v.max()) if v.size else LOWEST


def iterate(kind, face_mask, tol, start=0, end=None, flags=None, **arrays):
    """Refinement::iterate(start, end, flagger) -> (flags, count).  flags=None: cleared first."""
    mask = np.asarray(face_mask) != 0
    n = mask.shape[0]
    end = n if end is None else end
    flags = np.zeros(n, dtype=np.uint8) if flags is None else flags
    v = flag_values(kind, face_mask, **arrays)
    hit = (~mask) & (v > tol)
    hit[:start] = False
    hit[end:] = False
    flags[hit] = 1
    return flags, int(np.count_nonzero(flags[start:end]))


def ref_lib():
    """oracle/_ref/liblpm_ref.so (the reference header compiled in place), or None where it has not been built."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liblpm_ref.so")
    if not os.path.exists(path):
        return None
    L = ctypes.CDLL(path)
    if not hasattr(L, "ref_flag_scalar"):
        return None
    return L


def ref_iterate(L, kind, face_mask, rtol, relative, start=0, end=None, face_vals=None, area=None, vert_vals=None,
                face_verts=None, vert_lag=None):
    """The compiled reference functor: constructor(rtol) [+ set_tol_from_relative_value()] + iterate -> (flags, count, tol)."""
    import ctypes
    mask = np.ascontiguousarray(face_mask, dtype=np.uint8)
    n = mask.shape[0]
    end = n if end is None else end
    flags = np.zeros(n, dtype=np.uint8)
    tol = ctypes.c_double()
    vp = ctypes.c_void_p

    def ptr(a, dt):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=dt)
        return a, vp(a.ctypes.data)
    if kind == "flow_map_variation":
        fv, pfv = ptr(face_verts, np.int32)
        vl, pvl = ptr(vert_lag, np.float64)
        L.ref_flag_flow_map.argtypes = [ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, vp, ctypes.c_double, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_double)]
        ct = L.ref_flag_flow_map(fv.shape[1], vl.shape[0], pvl, n, pfv, vp(mask.ctypes.data), float(rtol), int(relative),
                                 start, end, vp(flags.ctypes.data), ctypes.byref(tol))
    else:
        fvals, pf = ptr(face_vals, np.float64)
        ar, pa = ptr(area, np.float64)
        vv, pv = ptr(vert_vals, np.float64)
        fv, pfv = ptr(face_verts, np.int32)
        L.ref_flag_scalar.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, vp,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                      ctypes.POINTER(ctypes.c_double)]
        ct = L.ref_flag_scalar(KINDS.index(kind), n, pf, pa, 0 if vv is None else vv.shape[0], pv, pfv,
                               0 if fv is None else fv.shape[1], vp(mask.ctypes.data), float(rtol), int(relative), start, end,
                               vp(flags.ctypes.data), ctypes.byref(tol))
    return flags, ct, tol.value
