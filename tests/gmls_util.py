"""Helpers of the GMLS tests (test infrastructure): ctypes access to the host build of the product's per-target
arithmetic (oracle/_build/libgmls_core_host.so, compiled from lpm_b200/csrc/lpmx_gmls_core.h) and real spherical
harmonics for the analytic anchor lap Y_l^m = -l(l+1) Y_l^m."""
import ctypes
import os
import subprocess

import numpy as np

_ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_LIB = os.path.join(_ORACLE_DIR, "_build", "libgmls_core_host.so")
_dp = ctypes.POINTER(ctypes.c_double)


def core_host_laplacian(xyz, f, p, radius=1.0):
    """(lap, eps, n_neighbors) from lpmx::gmls::laplacian_at_target run on the CPU."""
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-C", _ORACLE_DIR], check=True, capture_output=True)
    L = ctypes.CDLL(_LIB)
    xyz, f = np.ascontiguousarray(xyz, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64)
    n = f.shape[0]
    lap, eps, nn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rc = L.gmls_core_host_laplacian(n, xyz.ctypes.data_as(_dp), f.ctypes.data_as(_dp), p["samples_order"], p["manifold_order"],
                                    p["min_neighbors"], ctypes.c_double(p["eps_multiplier"]), ctypes.c_double(p["weight_pwr"]),
                                    ctypes.c_double(radius), lap.ctypes.data_as(_dp), eps.ctypes.data_as(_dp),
                                    nn.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0
    return lap, eps, nn


def core_host_interpolate(src_xyz, src_fields, tgt_xyz, p, radius=1.0):
    """(n_fields, n_tgt) values from lpmx::gmls::interpolate_at_point run on the CPU."""
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-C", _ORACLE_DIR], check=True, capture_output=True)
    L = ctypes.CDLL(_LIB)
    src = np.ascontiguousarray(src_xyz, dtype=np.float64)
    F = np.ascontiguousarray(np.atleast_2d(src_fields), dtype=np.float64)
    tgt = np.ascontiguousarray(tgt_xyz, dtype=np.float64)
    out = np.zeros((F.shape[0], tgt.shape[0]))
    rc = L.gmls_core_host_interpolate(src.shape[0], src.ctypes.data_as(_dp), F.shape[0], F.ctypes.data_as(_dp), tgt.shape[0],
                                      tgt.ctypes.data_as(_dp), p["samples_order"], p["min_neighbors"],
                                      ctypes.c_double(p["eps_multiplier"]), ctypes.c_double(p["weight_pwr"]),
                                      ctypes.c_double(radius), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


def remesh_case(depth, amp=0.05):
    """sources: an advected cubed-sphere particle set; targets: a fresh icosahedral one; five smooth fields + exact values"""
    from lpm_b200.api import PolyMesh2d
    from oracle import gmls_oracle as GO
    m = PolyMesh2d("cubed", depth)
    src = GO.gather(m.vert_xyz, m.face_xyz, m.face_mask)
    src = src + amp * np.stack([np.sin(2 * src[:, 1]), src[:, 2] * src[:, 0], np.cos(3 * src[:, 0]) * src[:, 1]], axis=1)
    src /= np.linalg.norm(src, axis=1)[:, None]
    mt = PolyMesh2d("icos", depth)
    tgt = GO.gather(mt.vert_xyz, mt.face_xyz, mt.face_mask)

    def fields(x):
        return np.stack([harmonic_field(x)[0], x[:, 0], x[:, 1], x[:, 2], real_sph_harm(x, 3, 2)])
    return np.ascontiguousarray(src), fields(src), np.ascontiguousarray(tgt), fields(tgt)


def real_sph_harm(xyz, l, m):
    from scipy.special import sph_harm_y
    r = np.linalg.norm(xyz, axis=1)
    theta = np.arccos(np.clip(xyz[:, 2] / r, -1, 1))
    phi = np.arctan2(xyz[:, 1], xyz[:, 0])
    return sph_harm_y(l, m, theta, phi).real


def harmonic_field(xyz):
    """f and its exact Laplace-Beltrami on the unit sphere"""
    y43, y21 = real_sph_harm(xyz, 4, 3), real_sph_harm(xyz, 2, 1)
    return y43 + 0.5 * y21, -20.0 * y43 - 3.0 * y21
