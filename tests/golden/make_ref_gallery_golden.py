"""Regenerate tests/golden/ref_gallery.npz: the REFERENCE's initial-condition functors (vorticity / velocity / surface gallery,
Coriolis, atan4, the Lamb dipole's Bessel functions), compiled in place (oracle/ref_gallery_driver.cpp ->
oracle/_ref/liblpm_ref.so), evaluated on seeded points.  Run in the build container after `make -C oracle ref`:
    python tests/golden/make_ref_gallery_golden.py
Points: 400 on the unit sphere (incl. the poles, the axes and points on the x-z / y-z planes, where atan4 branches) and 400 in
the plane (incl. both dipole centres, the dipole rims and the origin)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

SCALAR_IDS = {"solid_body_rotation": (0, 3, 0.0, 0.0), "gaussian_vortex": (1, 3, 0.0, 0.0), "gaussian_vortex_gc": (1, 3, 0.37, 0.0),
              "rh54": (2, 3, 0.0, 1.0), "rh54_stationary": (2, 3, 2 * np.pi / 14, 1.0), "rh54_scaled": (2, 3, 0.3, 0.25),
              "tc2_vorticity": (3, 3, 0.0, 0.0), "tc2_surface": (4, 3, 0.0, 0.0), "coriolis_sphere": (5, 3, 2 * np.pi, 0.0),
              "gaussian_mountain": (10, 2, 0.0, 0.0), "gaussian_mountain_laplacian": (11, 2, 0.0, 0.0),
              "surface_perturbation": (12, 2, 0.0, 0.0), "colliding_dipoles": (13, 2, 0.0, 0.0),
              "coriolis_beta_plane": (14, 2, 0.7, 0.2)}


def points():
    rng = np.random.default_rng(20261017)
    s = rng.standard_normal((380, 3))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    special = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]
    for a in np.linspace(0.1, 3.0, 7):  # on the coordinate planes: x = 0 or y = 0 exactly
        special += [[np.sin(a), 0.0, np.cos(a)], [-np.sin(a), 0.0, np.cos(a)]]
    sph = np.concatenate([np.array(special, dtype=np.float64), s])[:400]
    p = rng.uniform(-3.0, 3.0, (380, 2))
    th = np.linspace(0, 2 * np.pi, 8, endpoint=False)
    ring = np.stack([-1.5 + 0.999 * np.cos(th), 0.999 * np.sin(th)], axis=1)
    pl = np.concatenate([np.array([[0.0, 0.0], [-1.5, 0.0], [1.5, 0.0], [-1.5, 0.5], [1.5, -0.5], [-0.5, 0.0], [2.5, 0.0],
                                   [1.5, 1.0], [-1.5, -1.0], [-1.125, 0.0], [0.3, -0.2], [1.0, 1.0]]), ring, p])[:400]
    return np.ascontiguousarray(sph), np.ascontiguousarray(pl)


if __name__ == "__main__":
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblpm_ref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    L.ref_gallery_scalar.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, ctypes.c_double, dp]
    L.ref_gallery_rh54_velocity.argtypes = [ctypes.c_int, dp, ctypes.c_double, ctypes.c_double, dp]
    L.ref_gallery_rh54_stationary_u0.restype = ctypes.c_double
    L.ref_gallery_rh54_stationary_u0.argtypes = [ctypes.c_double]
    L.ref_bessel_j.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp]
    L.ref_atan4.argtypes = [ctypes.c_int, dp, dp, dp]
    sph, pl = points()
    out = {"sphere_points": sph, "plane_points": pl}
    for name, (fid, nd, p0, p1) in SCALAR_IDS.items():
        pts = sph if nd == 3 else pl
        v = np.zeros(len(pts))
        assert L.ref_gallery_scalar(fid, len(pts), pts.ctypes.data_as(dp), p0, p1, v.ctypes.data_as(dp)) == 0
        out[name] = v
        out[name + "_params"] = np.array([p0, p1])
    u = np.zeros((len(sph), 3))
    L.ref_gallery_rh54_velocity(len(sph), sph.ctypes.data_as(dp), 2 * np.pi / 14, 1.0, u.ctypes.data_as(dp))
    out["rh54_velocity_stationary"] = u
    out["rh54_stationary_u0"] = np.array([L.ref_gallery_rh54_stationary_u0(2 * np.pi)])
    xb = np.concatenate([np.linspace(0.0, 12.0, 481), [3.8317, 1e-9, 8.0, 7.999, 8.001]])
    for order in (0, 1):
        b = np.zeros(len(xb))
        L.ref_bessel_j(order, len(xb), xb.ctypes.data_as(dp), b.ctypes.data_as(dp))
        out[f"bessel_j{order}"] = b
    out["bessel_x"] = xb
    ya, xa = np.ascontiguousarray(sph[:, 1]), np.ascontiguousarray(sph[:, 0])
    a4 = np.zeros(len(sph))
    L.ref_atan4(len(sph), ya.ctypes.data_as(dp), xa.ctypes.data_as(dp), a4.ctypes.data_as(dp))
    out["atan4"] = a4
    np.savez_compressed(os.path.join(HERE, "ref_gallery.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items() if not k.endswith("_params")})
