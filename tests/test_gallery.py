"""CPU: the initial conditions are part of the parity claim ("the same mesh seeds and initial conditions").
tests/golden/ref_gallery.npz holds the REFERENCE's functors (src/lpm_vorticity_gallery.hpp, lpm_velocity_gallery.hpp,
lpm_surface_gallery.hpp, lpm_coriolis.hpp, util/lpm_math.hpp) compiled in place and evaluated on seeded points
(tests/golden/make_ref_gallery_golden.py).  Checked against it:
  * lpm_b200/gallery.py and tests/plane_cases.py -- what bench.py and the parity tests feed the engine;
  * the C++ shim's gallery (include/lpm/lpm_gallery.hpp, lpm_plane.hpp) -- what the example drivers use -- through a host-only
    program;
  * the live reference build when present.
Tolerance 2e-15 relative to the field's maximum: the expressions are the reference's, the differences are libm / FMA-contraction
round-off.  The Python RH54 gets 6e-15: numpy's arctan2 and glibc's atan2 differ by an ulp of the longitude, which cos(4 lon) times the
amplitude 30 turns into 2.5e-14 absolute.  The Lamb dipoles are the one flagged deviation: the shim evaluates J0, J1 with std::cyl_bessel_j, the reference with
its own rational approximations, which are themselves only good to ~5e-8 (their J0(0) is 1 + 2.8e-9)."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import plane_cases  # noqa: E402
from lpm_b200 import gallery  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "ref_gallery.npz"))
TOL = 2e-15
TOL_PY_RH54 = 6e-15
ORDER = ["solid_body_rotation", "gaussian_vortex", "gaussian_vortex_gc", "rh54", "rh54_stationary", "rh54_scaled", "tc2_vorticity",
         "tc2_surface", "coriolis_sphere", "gaussian_mountain", "gaussian_mountain_laplacian", "surface_perturbation",
         "colliding_dipoles", "coriolis_beta_plane"]
PLANAR = {"gaussian_mountain", "gaussian_mountain_laplacian", "surface_perturbation", "colliding_dipoles", "coriolis_beta_plane"}


def rel(a, b):
    return np.abs(np.asarray(a) - b).max() / np.abs(b).max()


def test_python_gallery_matches_the_reference_functors():
    x = G["sphere_points"]
    assert rel(gallery.SolidBodyRotation()(x), G["solid_body_rotation"]) <= TOL
    gv = gallery.GaussianVortexSphere()
    assert rel(gv(x), G["gaussian_vortex"]) <= TOL
    gv.set_gauss_const(0.37)
    assert rel(gv(x), G["gaussian_vortex_gc"]) <= TOL
    assert rel(gallery.RossbyHaurwitz54(0.0, 1.0)(x), G["rh54"]) <= TOL_PY_RH54
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    assert f.u0 == G["rh54_stationary_u0"][0]
    assert rel(f(x), G["rh54_stationary"]) <= TOL_PY_RH54
    assert rel(gallery.RossbyHaurwitz54(0.3, 0.25)(x), G["rh54_scaled"]) <= TOL_PY_RH54
    tc = gallery.SphereTestCase2()
    assert rel(tc.vorticity(x), G["tc2_vorticity"]) <= TOL
    assert rel(tc.surface(x), G["tc2_surface"]) <= TOL  # SphereTestCase2InitialSurface as coded (no u0^2/2 term)
    assert rel(2 * (2 * gallery.PI) * x[:, 2], G["coriolis_sphere"]) <= TOL
    assert np.abs(gallery.atan4(x[:, 1], x[:, 0]) - G["atan4"]).max() <= 4e-16 * 2 * np.pi
    assert gallery.PI == np.pi


def test_planar_test_fields_match_the_reference_functors():
    p = G["plane_points"]
    assert rel(plane_cases.gaussian_mountain(p), G["gaussian_mountain"]) <= TOL
    assert rel(plane_cases.surface_perturbation(p), G["surface_perturbation"]) <= TOL


@pytest.fixture(scope="module")
def shim_values(tmp_path_factory):
    d = tmp_path_factory.mktemp("gallery")
    exe = str(d / "gallery_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "gallery_check.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "lpm_b200"), "-llpmx", "-Wl,-rpath," + os.path.join(ROOT, "lpm_b200")],
                   check=True, capture_output=True)
    s, p = G["sphere_points"], G["plane_points"]
    s.tofile(str(d / "s.bin"))
    p.tofile(str(d / "p.bin"))
    subprocess.run([exe, str(d / "s.bin"), str(len(s)), str(d / "p.bin"), str(len(p)), str(d / "out.bin")], check=True)
    v = np.fromfile(str(d / "out.bin"))
    out, o = {}, 0
    for name in ORDER:
        n = len(p) if name in PLANAR else len(s)
        out[name] = v[o:o + n]
        o += n
    out["rh54_velocity_stationary"] = v[o:o + 3 * len(s)].reshape(-1, 3)
    o += 3 * len(s)
    out["atan4"] = v[o:o + len(s)]
    assert o + len(s) == len(v)
    return out


@pytest.mark.parametrize("name", [n for n in ORDER if n != "colliding_dipoles"] + ["rh54_velocity_stationary"])
def test_shim_gallery_matches_the_reference_functors(shim_values, name):
    assert rel(shim_values[name], G[name]) <= TOL, name


def test_shim_atan4_matches(shim_values):
    assert np.abs(shim_values["atan4"] - G["atan4"]).max() <= 4e-16 * 2 * np.pi


def test_shim_lamb_dipoles_agree_to_the_accuracy_of_the_reference_bessel_functions(shim_values):
    """Flagged deviation: std::cyl_bessel_j against the reference's rational approximations of J0 / J1."""
    from scipy.special import j0, j1
    x = G["bessel_x"]
    lo = x < 8.0  # the dipole needs k r <= 3.8317; the reference's bessel_j1 is off by up to 0.54 on its x >= 8 branch (not used)
    e0, e1 = np.abs(G["bessel_j0"] - j0(x))[lo].max(), np.abs(G["bessel_j1"] - j1(x))[lo].max()
    assert 1e-9 < e0 < 1e-8 and 1e-9 < e1 < 1e-8  # the reference's own error level: ~5e-9
    assert np.abs(G["bessel_j1"] - j1(x))[~lo].max() > 0.1
    assert rel(shim_values["colliding_dipoles"], G["colliding_dipoles"]) < 2e-7
    # support and sign structure are identical: zero outside both discs, at the centres and on the x axis (sin(theta) = y / r)
    assert np.array_equal(shim_values["colliding_dipoles"] == 0.0, G["colliding_dipoles"] == 0.0)


def test_live_reference_build_reproduces_the_golden():
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "liblpm_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = ctypes.CDLL(path)
    if not hasattr(L, "ref_gallery_scalar"):
        pytest.skip("oracle/_ref predates the gallery driver")
    dp = ctypes.POINTER(ctypes.c_double)
    L.ref_gallery_scalar.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, ctypes.c_double, dp]
    x = np.ascontiguousarray(G["sphere_points"])
    v = np.zeros(len(x))
    assert L.ref_gallery_scalar(2, len(x), x.ctypes.data_as(dp), 2 * np.pi / 14, 1.0, v.ctypes.data_as(dp)) == 0
    assert np.array_equal(v, G["rh54_stationary"])
