"""CPU: pin the oracle (oracle/lpm_oracle.c) before trusting it.

  1. the reference's own known-answer vectors (tests/lpm_swe_kernels_tests.cpp:47-50, 90-99, 122-127);
  2. golden outputs of the reference's own functors compiled in place (tests/golden/ref_sums.npz, made by
     tests/golden/make_ref_golden.py from oracle/_ref) -- and, where oracle/_ref exists, a live comparison;
  3. analytic solutions (solid-body rotation, examples/bve_rotation.cpp:147-167);
  4. the documented quirks of the reference.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import field_rel_err
from lpm_b200 import gallery
from lpm_b200.api import PolyMesh2d

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PI = np.pi


def _pt(lon, lat):
    return np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])


KZETA_EXACT = np.array([0.11761244724273439212, -0.30509169352258549181, 0.32004709265607662522])
KSIGMA_EXACT = np.array([-0.32583437560369497620, 0.16407250414068629817, 0.27614478771192878158])
# the value grad_kzeta() actually returns: the commented-out vector at :90-93; the live vector at :94-99 is
# its negative, which is presumably why the reference disables this test (tests/CMakeLists.txt:75)
GRAD_KZETA_AS_CODED = np.array([0.40473666151212280247, -0.53131662597042255196, 0.21608551568911567310,
                                -0.83791439363248985370, 0.82834755502308796106, 0.016331615064627251330,
                                0.43214450841002140895, 0.29017950208810278985, -1.2330842165352107635])
GRAD_KSIGMA_EXACT = np.array([0.39010639734131174212, -0.020240461517800895467, -0.63140064717422376023,
                              -0.34028755417387752068, -0.28570943666391770862, 1.0686437080615732960,
                              -0.93649234069680925204, 0.95103126081883890383, -0.024819489131446365611])


def test_reference_known_answer_vectors(oracle):
    x, y = _pt(PI / 4, PI / 8), _pt(PI / 6, PI / 20)
    kz, ks, gkz, gks = oracle.swe_pair(x, y, 0.0)
    assert np.abs(kz - KZETA_EXACT).max() < 1e-15
    assert np.abs(ks - KSIGMA_EXACT).max() < 1e-15
    assert np.abs(gkz - GRAD_KZETA_AS_CODED).max() < 1e-14
    assert np.abs(gks - GRAD_KSIGMA_EXACT).max() < 1e-14
    assert abs(kz @ x) < 2.3e-16 and abs(ks @ x) < 2.3e-16  # tangency (:85-88)
    kz1, ks1, _, _ = oracle.swe_pair(x, y, 0.01)  # eps = 0.01 within 20 eps^2 (:154,156)
    assert np.allclose(kz1, KZETA_EXACT, rtol=20 * 0.01 ** 2) and np.allclose(ks1, KSIGMA_EXACT, rtol=20 * 0.01 ** 2)
    # the same vector pins biot_savart and the IC2D velocity kernel (unit zeta*A)
    u = oracle.bve_velocity(x[None], y[None], [1.0], [1.0], [0])
    assert np.abs(u[0] - KZETA_EXACT).max() < 1e-15
    v, _ = oracle.ic2d_sums(x[None], y[None], [1.0], [1.0], [0], eps=0.0)
    assert np.abs(v[0] - KZETA_EXACT).max() < 1e-15


def test_pair_level_values_match_reference_code(oracle):
    g = np.load(os.path.join(GOLDEN, "ref_sums.npz"))
    worst = 0.0
    for x, y, eps, val in zip(g["pair_x"], g["pair_y"], g["pair_eps"], g["pair_vals"]):
        got = np.concatenate(oracle.swe_pair(x, y, float(eps)))
        for sl in (slice(0, 3), slice(3, 6), slice(6, 15), slice(15, 24)):
            worst = max(worst, np.abs(got[sl] - val[sl]).max() / np.abs(val[sl]).max())
    assert worst < 5e-14  # closed-form gradients vs the reference's expanded polynomials


def _sum_cases(o, L=None):
    out = {}
    for seed, depth in (("icos", 2), ("cubed", 3)):
        m = PolyMesh2d(seed, depth)
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        fz = f(m.face_xyz)
        sig = 0.3 * m.face_xyz[:, 0] * m.face_xyz[:, 2]
        k = f"{seed}{depth}_"
        a = (m.face_xyz, fz, m.face_area, m.face_mask)
        out[k + "bve_vel_verts"] = o.bve_velocity(m.vert_xyz, *a, L=L)
        out[k + "bve_vel_faces"] = o.bve_velocity(None, *a, collocated=True, L=L)
        out[k + "bve_psi_verts"] = o.bve_streamfn(m.vert_xyz, *a, L=L)
        out[k + "bve_psi_faces"] = o.bve_streamfn(None, *a, collocated=True, L=L)
        for eps in (0.0, 0.05):
            e = f"eps{eps}_"
            out[k + e + "ic2d_vel_passive"], out[k + e + "ic2d_psi_passive"] = o.ic2d_sums(m.vert_xyz, *a, eps=eps, L=L)
            out[k + e + "ic2d_vel_active"], out[k + e + "ic2d_psi_active"] = o.ic2d_sums(
                None, *a, eps=eps, targets_are_sources=True, L=L)
            out[k + e + "swe_ddot_verts"] = o.swe_sphere_sums(m.vert_xyz, m.face_xyz, fz, sig, m.face_area,
                                                              m.face_mask, eps=eps, L=L)[1]
            out[k + e + "swe_ddot_faces"] = o.swe_sphere_sums(None, m.face_xyz, fz, sig, m.face_area, m.face_mask,
                                                              eps=eps, targets_are_sources=True, L=L)[1]
    return out


_LEAF = {}


def _regular(name, n):
    """Targets on which the reference is well conditioned: everything, except that on the icosahedral mesh
    the divided faces of a singular (eps = 0) face sum sit on top of a leaf (d ~ 1e-16: NaN or O(1) noise)."""
    if name.startswith("icos2") and ("faces" in name or "active" in name) and ("eps0.05" not in name):
        if "icos2" not in _LEAF:
            _LEAF["icos2"] = PolyMesh2d("icos", 2).face_mask == 0
        return _LEAF["icos2"]
    return np.ones(n, dtype=bool)


def _compare(got, ref):
    for name, b in ref.items():
        a = got[name]
        fin = np.isfinite(b) if b.ndim == 1 else np.isfinite(b).all(axis=1)
        # NaN pattern of the reference (divided icos faces at eps = 0) is reproduced
        fa = np.isfinite(a) if a.ndim == 1 else np.isfinite(a).all(axis=1)
        assert np.array_equal(fin, fa), name
        reg = _regular(name, len(b))
        assert fin[reg].all(), name
        if name.split("_", 1)[1].startswith("bve"):
            assert np.array_equal(a[fin], b[fin]), name  # same operations in the same order: bit-identical
        elif "swe_ddot" in name:
            # quadratic in the 9 gradient sums; the reference's ~900-term expanded polynomial carries its own
            # ~1e-14 rounding noise per pair relative to the closed form
            assert field_rel_err(a, b, reg) < 1e-12, name
        else:
            assert field_rel_err(a, b, reg) < 5e-14, name


def test_sums_match_golden_reference_outputs(oracle):
    g = np.load(os.path.join(GOLDEN, "ref_sums.npz"))
    ref = {k: g[k] for k in g.files if k[:4] in ("icos", "cube") and not k.endswith(("zeta", "sigma"))}
    _compare(_sum_cases(oracle), ref)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "liblpm_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_sums_match_live_reference_build(oracle):
    R = ctypes.CDLL(oracle.REF_LIB)
    _compare(_sum_cases(oracle), _sum_cases(oracle, L=R))


def _tend_case(o, g, L=None):
    out = []
    for is_area in (0, 1):
        out.append(np.stack(o.swe_tendencies(is_area, g["tend_in_x"], g["tend_in_u"], g["tend_in_zeta"],
                                             g["tend_in_sigma"], g["tend_in_third"], g["tend_in_ddot"],
                                             g["tend_in_laps"], Omega=2 * np.pi, g=1.5, dt=0.0125, L=L)))
    return out


def test_swe_tendencies_match_golden_reference_outputs(oracle):
    """SWEVorticityDivergence{Height,Area}Tendencies (lpm_swe_kernels.hpp:941-1069) as compiled from the reference."""
    g = np.load(os.path.join(GOLDEN, "ref_sums.npz"))
    got = _tend_case(oracle, g)
    for is_area in (0, 1):
        ref = g[f"tend_out_{is_area}"]
        assert np.abs(got[is_area] - ref).max() <= 4e-15 * np.abs(ref).max()  # FMA contraction may differ by an ulp


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "liblpm_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_swe_surface_setters_match_live_reference_build(oracle):
    """SetSurfaceFromDepth / SetDepthAndSurfaceFromMassAndArea with ZeroFunctor (lpm_swe_kernels.hpp:1079-1140)."""
    R = ctypes.CDLL(oracle.REF_LIB)
    Lo = oracle.lib()
    rng = np.random.default_rng(5)
    n = 100
    dp = ctypes.POINTER(ctypes.c_double)
    x = rng.standard_normal((n, 3))
    h, m, area = 1 + rng.random(n), 1 + rng.random(n), 0.1 + rng.random(n)
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    s1, b1, s2, b2 = (np.full(n, 7.0) for _ in range(4))
    R.ref_swe_set_surface_from_depth(n, P(s1), P(b1), P(x), P(h))
    Lo.oracle_swe_set_surface_from_depth(n, P(s2), P(b2), P(h))
    assert np.array_equal(s1, s2) and np.array_equal(b1, b2)
    out1 = [np.full(n, 7.0) for _ in range(3)]
    out2 = [np.full(n, 7.0) for _ in range(3)]
    mp = mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte))
    R.ref_swe_set_depth_surface_from_mass_area(n, *map(P, out1), P(x), P(m), P(area), mp)
    Lo.oracle_swe_set_depth_surface_from_mass_area(n, *map(P, out2), P(m), P(area), mp)
    for a, b in zip(out1, out2):
        assert np.array_equal(a, b)
    assert (out1[0][mask == 1] == 7.0).all()  # divided faces are left untouched


def test_solid_body_rotation_analytic(oracle):
    """examples/bve_rotation.cpp:147-167: zeta = 2 Omega z  ->  u = Omega(-y, x, 0), psi = Omega z; the direct
    sums converge to it at first order in the mesh size."""
    sbr = gallery.SolidBodyRotation()
    errs = []
    for depth in (2, 3, 4):
        m = PolyMesh2d("icos", depth)
        fz = sbr(m.face_xyz)
        u = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        p = oracle.bve_streamfn(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        errs.append((np.abs(u - sbr.velocity(m.vert_xyz)).max(), np.abs(p - sbr.stream_fn(m.vert_xyz)).max()))
    errs = np.array(errs)
    assert (errs[1:, 0] < 0.75 * errs[:-1, 0]).all(), errs  # velocity: singular kernel, ~ h^0.6 in max-norm
    assert (errs[1:, 1] < 0.30 * errs[:-1, 1]).all(), errs  # stream function: ~ h^2
    assert errs[-1, 0] < 0.02 and errs[-1, 1] < 0.003


def test_quirk_divided_icos_faces_are_degenerate_targets(oracle):
    """SURVEY.md A-iii: all faces are targets.  A divided icosahedral face's centre coincides with its centre
    descendant's (d = 1 - x.y <= 2.3e-16), so the reference returns NaN there; leaves are unaffected."""
    m = PolyMesh2d("icos", 3)
    fz = gallery.SolidBodyRotation()(m.face_xyz)
    u = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    leaf = m.face_mask == 0
    assert np.isfinite(u[leaf]).all()
    assert (~np.isfinite(u[~leaf]).all(axis=1)).sum() > 0


def test_quirk_rk4_face_vorticity_uses_k4_twice(oracle):
    """lpm_bve_rk4_impl.hpp:155-157: faces get zeta += (k1+k4)/6 + (k2+k4)/3; vertices the textbook formula.
    A vertex and a face at the same place with the same velocity history therefore end with different zeta."""
    m = PolyMesh2d("cubed", 2)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    a = (m.face_xyz, fz, m.face_area, m.face_mask)
    vu = oracle.bve_velocity(m.vert_xyz, *a)
    fu = oracle.bve_velocity(None, *a, collocated=True)
    # cubed sphere: a divided face's centre IS a vertex (lpm_faces_impl.hpp:506-517)
    parent = int(np.nonzero(m.face_mask == 1)[0][-1])
    vtx = int(np.nonzero((np.abs(m.vert_xyz - m.face_xyz[parent]).max(axis=1) == 0))[0][0])
    st = [m.vert_xyz.copy(), vz.copy(), vu, m.face_xyz.copy(), fz.copy(), fu]
    assert np.array_equal(st[2][vtx], st[5][parent])  # same position, same sources -> same velocity
    oracle.bve_rk4_step(0.05, 2 * PI, *st, m.face_area, m.face_mask, n_steps=1)
    assert np.array_equal(st[0][vtx], st[3][parent])   # positions advance identically
    assert st[1][vtx] != st[4][parent]                # vorticity does not: the quirk
    assert abs(st[1][vtx] - st[4][parent]) < 1e-2  # = (k4 - k3) / 3, an O(dt^2) effect


def test_ic2d_self_term_rules(oracle):
    """SURVEY.md B-ii: with eps > 0 the self term is included for active sums: it adds 0 to u and
    -log(eps^2) * zeta A / (4 pi) to psi; with eps = 0 it is skipped by index."""
    m = PolyMesh2d("cubed", 2)
    fz = np.ones(m.n_faces)
    eps = 0.1
    _, psi = oracle.ic2d_sums(None, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps, targets_are_sources=True)
    _, psi_p = oracle.ic2d_sums(m.face_xyz, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps)  # as passive: same
    assert np.array_equal(psi, psi_p)
    leaf = np.nonzero(m.face_mask == 0)[0]
    i = leaf[0]
    area = m.face_area.copy()
    area[i] = 0.0  # remove source i
    _, psi_wo = oracle.ic2d_sums(None, m.face_xyz, fz, area, m.face_mask, eps=eps, targets_are_sources=True)
    expect = -np.log(1 - m.face_xyz[i] @ m.face_xyz[i] + eps ** 2) * m.face_area[i] / (4 * PI)
    assert abs((psi[i] - psi_wo[i]) - expect) < 1e-14


def test_long_double_adjudicator_agrees_to_roundoff(oracle):
    m = PolyMesh2d("cubed", 4)
    fz = gallery.SolidBodyRotation()(m.face_xyz)
    a = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
    b = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask, long_double=True)
    assert field_rel_err(a, b) < 1e-13


# ---- the stepper itself: BVESphere + BVERK4::advance_timestep of the reference, compiled in place -----------------------------
REF_RK4_CASES = ["icos3_rh54", "cubed3_rh54", "icos4_rot_3", "icos4_rot_100"]


def ref_rk4_case(name):
    """(seed, depth, Omega, dt, n_steps, golden dict without the prefix) of tests/golden/ref_bve_rk4.npz
    (tests/golden/make_ref_stepper_golden.py: outputs of oracle/_ref/liblpm_ref_mesh.so)."""
    g = np.load(os.path.join(GOLDEN, "ref_bve_rk4.npz"))
    depth, omega, dt, n_steps = g[f"{name}_params"]
    d = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_")}
    return name.split("_")[0][:-1], int(depth), float(omega), float(dt), int(n_steps), d


@pytest.mark.parametrize("name", REF_RK4_CASES)
def test_oracle_stepper_matches_compiled_reference_bve_rk4(oracle, name):
    """oracle_bve_rk4_step (the restatement every GPU stepper test is checked against) vs the reference's own
    BVERK4::advance_timestep (src/lpm_bve_rk4_impl.hpp:63-167: 4 x {BVEVertexVelocity, BVEFaceVelocity, BVEVorticityTendency},
    20 KokkosBlas calls, the facevort4-twice update) after BVESphere::init_velocity, incl. 100 steps at icos-4."""
    seed, depth, omega, dt, n_steps, g = ref_rk4_case(name)
    m = PolyMesh2d(seed, depth)
    vz, fz = g["vert_zeta0"], g["face_zeta0"]
    vu = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
    fu = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    st = [m.vert_xyz.copy(), vz.copy(), vu, m.face_xyz.copy(), fz.copy(), fu]
    oracle.bve_rk4_step(dt, omega, *st, m.face_area, m.face_mask, n_steps=n_steps)
    leaf = m.face_mask == 0
    tol = 2e-14 if n_steps <= 3 else 2e-13  # round-off of two compilations of the same arithmetic (FMA contraction differs)
    assert field_rel_err(st[0], g["vert_xyz"]) <= tol
    assert field_rel_err(st[3], g["face_xyz"], leaf) <= tol
    assert field_rel_err(st[1], g["vert_zeta"]) <= tol
    assert field_rel_err(st[4], g["face_zeta"], leaf) <= tol
    if "vert_vel" in g:
        assert field_rel_err(st[2], g["vert_vel"]) <= 10 * tol
        assert field_rel_err(st[5], g["face_vel"], leaf) <= 10 * tol
        psi_v = oracle.bve_streamfn(st[0], st[3], st[4], m.face_area, m.face_mask)
        psi_f = oracle.bve_streamfn(None, st[3], st[4], m.face_area, m.face_mask, collocated=True)
        assert field_rel_err(psi_v, g["vert_psi"]) <= 10 * tol
        assert field_rel_err(psi_f, g["face_psi"], leaf) <= 10 * tol


def test_rows_of_divided_icos_faces_depend_on_the_reference_build_flags(oracle):
    """Why every comparison on icosahedral meshes selects leaf rows: a divided TriFace keeps its centre, which is also the centre
    of its middle child, so 1 - x.y of that (target, source) pair is 0 or an ulp, and the reference's value there is 0/0 = NaN,
    or 1e16-sized, depending on whether the compiler contracted x.y into FMAs.  Shown on the reference itself: the same functor
    (BVEFaceVelocity, src/lpm_bve_sphere_kernels.hpp:284-320) compiled with contraction (oracle/_ref/liblpm_ref.so) and without
    (BVESphere::init_velocity in liblpm_ref_mesh.so) disagrees on WHICH rows are NaN, while all leaf rows agree to round-off.
    These rows are targets only (masked faces are never sources), so nothing observable depends on them."""
    from oracle import ref_mesh
    if not (os.path.exists(oracle.REF_LIB) and ref_mesh.available()):
        pytest.skip("oracle/_ref not built")
    L = ctypes.CDLL(oracle.REF_LIB)
    m = PolyMesh2d("icos", 2)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    a = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True, L=L)  # contracted build
    b = ref_mesh.bve_rk4_run("icos", 2, 0.01, 0.0, 0, vz, fz)["face_vel"]                          # -ffp-contract=off build
    leaf = m.face_mask == 0
    assert field_rel_err(a, b, leaf) <= 1e-14
    nan_a, nan_b = np.isnan(a).any(axis=1), np.isnan(b).any(axis=1)
    assert not nan_a[leaf].any() and not nan_b[leaf].any()
    assert nan_a.sum() > 0 and nan_b.sum() > 0 and not np.array_equal(nan_a, nan_b)


# ---- Incompressible2D + Incompressible2DRK2 of the reference, compiled in place ------------------------------------------------
REF_IC2D_CASES = ["cubed3_rh54", "icos3_rh54", "cubed4_gauss"]


def ref_ic2d_case(name):
    """(seed, depth, Omega, dt, n_steps, eps, golden dict) of tests/golden/ref_ic2d_rk2.npz (make_ref_stepper_golden.py: outputs of
    oracle/_ref/liblpm_ref_mesh.so, i.e. the reference's Incompressible2D<Seed>::init_direct_sums + n x advance_timestep)."""
    g = np.load(os.path.join(GOLDEN, "ref_ic2d_rk2.npz"))
    depth, omega, dt, n_steps, eps = g[f"{name}_params"]
    d = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_")}
    return name.split("_")[0][:-1], int(depth), float(omega), float(dt), int(n_steps), float(eps), d


@pytest.mark.parametrize("name", REF_IC2D_CASES)
def test_oracle_stepper_matches_compiled_reference_ic2d_rk2(oracle, name):
    """oracle_ic2d_rk2_step vs the reference's own Incompressible2DRK2::advance_timestep_impl
    (src/lpm_incompressible2d_rk2_impl.hpp:75-172: PassiveSums/ActiveSums x 2, Incompressible2DTendencies without dt, 12 KokkosBlas
    calls) after Incompressible2D::init_direct_sums, eps = 0 (self term skipped) and eps > 0 (kept)."""
    seed, depth, omega, dt, n_steps, eps, g = ref_ic2d_case(name)
    m = PolyMesh2d(seed, depth)
    vz, fz = g["vert_zeta0"], g["face_zeta0"]
    pu, pp = oracle.ic2d_sums(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps)
    au, ap = oracle.ic2d_sums(None, m.face_xyz, fz, m.face_area, m.face_mask, eps=eps, targets_are_sources=True)
    st = [m.vert_xyz.copy(), vz.copy(), pu, pp, m.face_xyz.copy(), fz.copy(), au, ap]
    oracle.ic2d_rk2_step(dt, omega, eps, *st, m.face_area, m.face_mask, n_steps=n_steps)
    leaf = m.face_mask == 0
    keys = ["vert_xyz", "vert_zeta", "vert_vel", "vert_psi", "face_xyz", "face_zeta", "face_vel", "face_psi"]
    for k, a in zip(keys, st):
        sel = leaf if k.startswith("face") else None
        assert field_rel_err(a, g[k], sel) <= (2e-13 if k.endswith(("vel", "psi")) else 2e-14), k
