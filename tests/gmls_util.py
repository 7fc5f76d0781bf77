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
