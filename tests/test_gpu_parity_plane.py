"""GPU parity of the planar paths through the C ABI against the CPU oracle
(oracle/lpm_oracle_plane.c, bit-identical to the reference's own planar functors: tests/test_oracle_plane.py) and
against the committed outputs of the reference build (tests/golden/ref_plane.npz).
Tolerances (north_star): <= 1e-12 field-relative on the direct sums, <= 1e-10 on the advected state after n steps."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import plane_cases  # noqa: E402
from conftest import field_rel_err  # noqa: E402
from lpm_b200 import api  # noqa: E402
from lpm_b200.api import LAYOUT_LEFT, LAYOUT_RIGHT  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUM_TOL = 1e-12
STEP_TOL = 1e-10


@pytest.fixture(scope="module")
def OP(oracle):
    from oracle import oracle_plane
    return oracle_plane


def _left(a):
    return np.ascontiguousarray(a.T)


@pytest.mark.parametrize("eps", [0.0, 0.05])
@pytest.mark.parametrize("layout", [LAYOUT_RIGHT, LAYOUT_LEFT])
def test_ic2d_plane_sums(engine, OP, eps, layout):
    P, A, mask, h = plane_cases.quad_case(n=24, radius=2.0)
    cv = (lambda a: a) if layout == LAYOUT_RIGHT else _left
    back = (lambda a: a) if layout == LAYOUT_RIGHT else (lambda a: a.T)
    u, psi = api.ic2d_plane_sums(engine, cv(P["xy"]), cv(A["xy"]), A["vort"], A["area"], mask, eps=eps, layout=layout)
    ur, psir = OP.ic2d_plane_sums(P["xy"], A["xy"], A["vort"], A["area"], mask, eps=eps)
    assert field_rel_err(back(u), ur) < SUM_TOL and field_rel_err(psi, psir) < SUM_TOL
    u, psi = api.ic2d_plane_sums(engine, None, cv(A["xy"]), A["vort"], A["area"], mask, eps=eps,
                                 targets_are_sources=True, layout=layout)
    ur, psir = OP.ic2d_plane_sums(None, A["xy"], A["vort"], A["area"], mask, eps=eps, targets_are_sources=True)
    assert field_rel_err(back(u), ur) < SUM_TOL and field_rel_err(psi, psir) < SUM_TOL
    # velocity only
    u2, none = api.ic2d_plane_sums(engine, None, cv(A["xy"]), A["vort"], A["area"], mask, eps=eps,
                                   targets_are_sources=True, with_psi=False, layout=layout)
    assert none is None and np.array_equal(u2, u)


@pytest.mark.parametrize("eps", [0.0, 0.05])
def test_swe_plane_sums(engine, OP, eps):
    P, A, mask, h = plane_cases.quad_case(n=24, radius=2.0)
    pse = plane_cases.pse_eps_of(h)
    got = api.swe_plane_sums(engine, P["xy"], P["surf"], A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse)
    ref = OP.swe_plane_sums(P["xy"], P["surf"], A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse)
    errs = {k: field_rel_err(got[k], ref[k]) for k in api.PLANE_SUM_FIELDS}
    assert max(errs.values()) < SUM_TOL, errs
    got = api.swe_plane_sums(engine, None, None, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse,
                             targets_are_sources=True)
    ref = OP.swe_plane_sums(None, None, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse,
                            targets_are_sources=True)
    errs = {k: field_rel_err(got[k], ref[k]) for k in api.PLANE_SUM_FIELDS}
    assert max(errs.values()) < SUM_TOL, errs
    # do_velocity = false leaves the velocity alone and changes nothing else
    g2 = api.swe_plane_sums(engine, None, None, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse,
                            targets_are_sources=True, do_velocity=False)
    assert g2["vel"] is None and np.array_equal(g2["ddot"], got["ddot"]) and np.array_equal(g2["laps"], got["laps"])


def test_against_committed_reference_outputs(engine):
    """The reference build's own outputs (tests/golden/ref_plane.npz), no oracle in between."""
    g = np.load(os.path.join(GOLDEN, "ref_plane.npz"))
    P, A, mask, h = plane_cases.quad_case(n=12, radius=2.0)
    pse = plane_cases.pse_eps_of(h)
    worst = 0.0
    for eps in (0.0, 0.05):
        e = f"eps{eps}_"
        u, psi = api.ic2d_plane_sums(engine, P["xy"], A["xy"], A["vort"], A["area"], mask, eps=eps)
        worst = max(worst, field_rel_err(u, g[e + "ic2d_vel_passive"]), field_rel_err(psi, g[e + "ic2d_psi_passive"]))
        u, psi = api.ic2d_plane_sums(engine, None, A["xy"], A["vort"], A["area"], mask, eps=eps, targets_are_sources=True)
        worst = max(worst, field_rel_err(u, g[e + "ic2d_vel_active"]), field_rel_err(psi, g[e + "ic2d_psi_active"]))
        r = api.swe_plane_sums(engine, P["xy"], P["surf"], A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse)
        worst = max(worst, max(field_rel_err(r[k], g[e + "swe_passive_" + k]) for k in api.PLANE_SUM_FIELDS))
        r = api.swe_plane_sums(engine, None, None, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], eps, pse,
                               targets_are_sources=True)
        worst = max(worst, max(field_rel_err(r[k], g[e + "swe_active_" + k]) for k in api.PLANE_SUM_FIELDS))
    assert worst < SUM_TOL, worst


def test_plane_sums_edge_cases(engine, OP):
    P, A, mask, h = plane_cases.quad_case(n=8, radius=1.0)
    pse = plane_cases.pse_eps_of(h)
    # every source masked: all sums are exactly zero
    allm = np.ones_like(mask)
    r = api.swe_plane_sums(engine, P["xy"], P["surf"], A["xy"], A["vort"], A["div"], A["area"], allm, A["surf"], 0.0, pse)
    assert all(np.abs(r[k]).max() == 0.0 for k in api.PLANE_SUM_FIELDS)
    u, psi = api.ic2d_plane_sums(engine, P["xy"], A["xy"], A["vort"], A["area"], allm)
    assert np.abs(u).max() == 0.0 and np.abs(psi).max() == 0.0
    # a single source, targets including the origin (the padding records must not leak a 0 * log(0))
    src = np.array([[0.3, -0.2]])
    tgt = np.array([[0.0, 0.0], [1.0, 1.0], [-2.0, 0.5]])
    u, psi = api.ic2d_plane_sums(engine, tgt, src, [2.0], [0.5], [0])
    ur, psir = OP.ic2d_plane_sums(tgt, src, [2.0], [0.5], [0])
    assert np.isfinite(u).all() and field_rel_err(u, ur) < SUM_TOL and field_rel_err(psi, psir) < SUM_TOL
    # ragged sizes around the 256-source chunk and the target-block sizes
    rng = np.random.default_rng(5)
    for n_src, n_tgt in ((255, 1), (256, 33), (257, 1025), (513, 2049)):
        sx, tx = rng.uniform(-2, 2, (n_src, 2)), rng.uniform(-2, 2, (n_tgt, 2))
        z, ar = rng.standard_normal(n_src), rng.random(n_src) * 0.01
        m = (rng.random(n_src) < 0.1).astype(np.uint8)
        u, psi = api.ic2d_plane_sums(engine, tx, sx, z, ar, m, eps=0.01)
        ur, psir = OP.ic2d_plane_sums(tx, sx, z, ar, m, eps=0.01)
        assert field_rel_err(u, ur) < SUM_TOL and field_rel_err(psi, psir) < SUM_TOL, (n_src, n_tgt)
        ss, ts, sg = rng.random(n_src), rng.random(n_tgt), rng.standard_normal(n_src)
        r = api.swe_plane_sums(engine, tx, ts, sx, z, sg, ar, m, ss, 0.02, 0.3)
        rr = OP.swe_plane_sums(tx, ts, sx, z, sg, ar, m, ss, 0.02, 0.3)
        errs = {k: field_rel_err(r[k], rr[k]) for k in api.PLANE_SUM_FIELDS}
        assert max(errs.values()) < SUM_TOL, (n_src, n_tgt, errs)
    # argument errors are reported, not crashed on
    with pytest.raises(Exception):
        api.swe_plane_sums(engine, P["xy"], P["surf"], A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], 0.0, -1.0)


def test_pse_and_log_ranges(engine, OP):
    """exp(-q) and log(a) over the full dynamic range the kernels see: neighbours at 1e-6 up to sources 50 kernel
    widths away (exp(-2500) underflows to 0 on both sides)."""
    rng = np.random.default_rng(11)
    n_src = 1024
    rad = 10.0 ** rng.uniform(-6, 1.2, n_src)
    ang = rng.uniform(0, 2 * np.pi, n_src)
    sx = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    tx = np.array([[0.0, 0.0], [1e-3, -1e-3], [3.0, 4.0]])
    z, sg, ar = rng.standard_normal(n_src), rng.standard_normal(n_src), rad ** 2 * 0.1
    ss, ts = rng.random(n_src), rng.random(3)
    m = np.zeros(n_src, np.uint8)
    for pse in (0.3, 1e-3):
        r = api.swe_plane_sums(engine, tx, ts, sx, z, sg, ar, m, ss, 0.0, pse)
        rr = OP.swe_plane_sums(tx, ts, sx, z, sg, ar, m, ss, 0.0, pse)
        errs = {k: field_rel_err(r[k], rr[k]) for k in api.PLANE_SUM_FIELDS}
        assert max(errs.values()) < SUM_TOL, (pse, errs)


@pytest.mark.parametrize("eps,beta", [(0.0, 0.0), (0.08, 0.4)])
def test_ic2d_plane_rk2_step(engine, OP, eps, beta):
    P, A, mask, h = plane_cases.quad_case(n=16, radius=2.0, topo=False)
    px, pz, ax, az = P["xy"].copy(), P["vort"].copy(), A["xy"].copy(), A["vort"].copy()
    pu, ppsi = OP.ic2d_plane_sums(px, ax, az, A["area"], mask, eps=eps)
    au, apsi = OP.ic2d_plane_sums(None, ax, az, A["area"], mask, eps=eps, targets_are_sources=True)
    g = [a.copy() for a in (px, pz, pu, ppsi, ax, az, au, apsi)]
    n_steps, dt = 3, 0.02
    OP.ic2d_plane_rk2_step(dt, 0.1, beta, eps, px, pz, pu, ppsi, ax, az, au, apsi, A["area"], mask, n_steps=n_steps)
    api.ic2d_plane_rk2_step(engine, dt, 0.1, beta, eps, *g, A["area"], mask, n_steps=n_steps)
    names = ("px", "pz", "pu", "ppsi", "ax", "az", "au", "apsi")
    errs = {n: field_rel_err(a, b) for n, a, b in zip(names, g, (px, pz, pu, ppsi, ax, az, au, apsi))}
    assert max(errs.values()) < STEP_TOL, errs
    assert field_rel_err(g[4], A["xy"]) > 1e-5  # the particles did move


def _swe_state(OP, n, eps, pse_of=plane_cases.pse_eps_of, topo=True):
    P, A, mask, h = plane_cases.quad_case(n=n, radius=2.0, topo=topo)
    st = OP.PlaneSWEState(P, A, mask)
    pse = pse_of(h)
    OP.swe_plane_init_direct_sums(st, eps, pse)
    return st, pse


@pytest.mark.parametrize("eps,topo,layout", [(0.0, 1, LAYOUT_RIGHT), (0.05, 0, LAYOUT_RIGHT), (0.05, 1, LAYOUT_LEFT)])
def test_swe_plane_rk4_step(engine, OP, eps, topo, layout):
    st, pse = _swe_state(OP, 16, eps, topo=bool(topo))
    dt, f0, beta, g, n_steps = 0.01, 0.5, 0.2, 1.0, 3
    cv = (lambda a: a.copy()) if layout == LAYOUT_RIGHT else (lambda a: _left(a) if a.ndim == 2 else a.copy())
    back = (lambda a: a) if layout == LAYOUT_RIGHT else (lambda a: a.T if a.ndim == 2 else a)
    gp = {k: cv(st.p[k]) for k in api.PLANE_PASSIVE_FIELDS}
    ga = {k: cv(st.a[k]) for k in api.PLANE_ACTIVE_FIELDS}
    OP.swe_plane_rk4_step(dt, f0, beta, g, eps, pse, topo, st, n_steps=n_steps)
    api.swe_plane_rk4_step(engine, dt, f0, beta, g, eps, pse, topo, gp, ga, st.mask, n_steps=n_steps, layout=layout)
    leaf = st.mask == 0
    errs = {}
    for k in api.PLANE_PASSIVE_FIELDS:
        errs["p_" + k] = field_rel_err(back(gp[k]), st.p[k])
    for k in api.PLANE_ACTIVE_FIELDS:
        # divided panels are targets too: every field of theirs is compared as well
        errs["a_" + k] = field_rel_err(back(ga[k]), st.a[k])
    assert max(errs.values()) < STEP_TOL, errs
    assert field_rel_err(back(ga["xy"])[leaf], plane_cases.quad_case(n=16, radius=2.0, topo=bool(topo))[1]["xy"][leaf]) > 1e-6


def test_swe_plane_resident_solver_matches_in_place(engine, OP):
    eps, topo = 0.05, 1
    st, pse = _swe_state(OP, 12, eps)
    dt, f0, beta, g = 0.01, 0.5, 0.2, 1.0
    gp = {k: st.p[k].copy() for k in api.PLANE_PASSIVE_FIELDS}
    ga = {k: st.a[k].copy() for k in api.PLANE_ACTIVE_FIELDS}
    # resident: set the raw fields, let the engine do init_direct_sums, advance 2 + 2 steps, read back
    sol = api.PlaneSWESolver(engine, gp["xy"].shape[0], ga["xy"].shape[0], eps, pse, topo)
    raw_p = {k: gp[k] for k in ("xy", "vort", "div", "depth", "surf", "bottom")}
    raw_a = {k: ga[k] for k in ("xy", "vort", "div", "area", "mass", "depth", "surf", "bottom")}
    sol.set_state(raw_p, raw_a, st.mask)
    sol.init_direct_sums(True)
    op = {k: np.empty_like(gp[k]) for k in api.PLANE_PASSIVE_FIELDS}
    oa = {k: np.empty_like(ga[k]) for k in api.PLANE_ACTIVE_FIELDS}
    sol.get_state(op, oa)
    e0 = max(max(field_rel_err(op[k], st.p[k]) for k in api.PLANE_PASSIVE_FIELDS),
             max(field_rel_err(oa[k], st.a[k]) for k in api.PLANE_ACTIVE_FIELDS))
    assert e0 < SUM_TOL, e0  # init_direct_sums == the oracle's
    gp = {k: op[k].copy() for k in api.PLANE_PASSIVE_FIELDS}  # the in-place path starts from the same bits
    ga = {k: oa[k].copy() for k in api.PLANE_ACTIVE_FIELDS}
    sol.advance(dt, f0, beta, g, 2)
    sol.advance(dt, f0, beta, g, 2)
    sol.get_state(op, oa)
    api.swe_plane_rk4_step(engine, dt, f0, beta, g, eps, pse, topo, gp, ga, st.mask, n_steps=4)
    for k in api.PLANE_PASSIVE_FIELDS:
        assert np.array_equal(op[k], gp[k]), k  # same kernels, same order: bit-identical
    for k in api.PLANE_ACTIVE_FIELDS:
        assert np.array_equal(oa[k], ga[k]), k
    local, glob = sol.interactions_per_eval()
    n_leaf = int((st.mask == 0).sum())
    assert glob == (gp["xy"].shape[0] + ga["xy"].shape[0]) * n_leaf and local == glob
    sol.close()


def test_plane_sums_large_subset(engine, OP):
    """65k sources: a 384-target subset against the oracle, plus linearity in the source strengths (a property that
    does not need the oracle at full size)."""
    P, A, mask, h = plane_cases.quad_case(n=256, radius=6.0, with_parents=False)
    pse = plane_cases.pse_eps_of(h)
    rng = np.random.default_rng(3)
    sel = rng.choice(P["xy"].shape[0], 384, replace=False)
    tx, ts = np.ascontiguousarray(P["xy"][sel]), np.ascontiguousarray(P["surf"][sel])
    r = api.swe_plane_sums(engine, tx, ts, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], 0.0, pse)
    rr = OP.swe_plane_sums(tx, ts, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], 0.0, pse)
    errs = {k: field_rel_err(r[k], rr[k]) for k in api.PLANE_SUM_FIELDS}
    assert max(errs.values()) < SUM_TOL, errs
    full = api.swe_plane_sums(engine, None, None, A["xy"], A["vort"], A["div"], A["area"], mask, A["surf"], 0.0, pse,
                              targets_are_sources=True)
    fz = api.swe_plane_sums(engine, None, None, A["xy"], A["vort"], 0 * A["div"], A["area"], mask, A["surf"], 0.0, pse,
                            targets_are_sources=True)
    fs = api.swe_plane_sums(engine, None, None, A["xy"], 0 * A["vort"], A["div"], A["area"], mask, A["surf"], 0.0, pse,
                            targets_are_sources=True)
    for k in ("vel", "du1dx1", "du1dx2", "du2dx1", "du2dx2"):
        assert field_rel_err(fz[k] + fs[k], full[k]) < 1e-13, k
    assert np.abs(fz["phi"]).max() == 0.0 and np.abs(fs["psi"]).max() == 0.0
    assert np.array_equal(fz["laps"], full["laps"])  # the PSE term does not depend on the strengths


@pytest.mark.parametrize("seed", ["quad_rect", "tri_hex"])
def test_swe_plane_rk4_on_the_generated_planar_meshes(engine, OP, seed):
    """examples/plane_gravity_wave.cpp end to end at depth 3, radius 6: the mesh from the product's generator (PolyMesh2d of
    QuadRectSeed / TriHexSeed, divided parents included as targets), the example's surface and mountain, SWE::init_direct_sums
    and 3 SWERK4 steps on the device against the oracle stepper.  Tolerance: STEP_TOL field-relative."""
    from lpm_b200.api import PolyMesh2d
    m = PolyMesh2d(seed, 3, radius=6.0)
    nv, nf = m.n_verts, m.n_faces
    vb, fb = plane_cases.gaussian_mountain(m.vert_xyz), plane_cases.gaussian_mountain(m.face_xyz)
    vs, fs = plane_cases.surface_perturbation(m.vert_xyz), plane_cases.surface_perturbation(m.face_xyz)
    P = {"xy": m.vert_xyz, "vort": np.zeros(nv), "div": np.zeros(nv), "depth": vs - vb, "surf": vs, "bottom": vb}
    A = {"xy": m.face_xyz, "vort": np.zeros(nf), "div": np.zeros(nf), "area": m.face_area, "mass": (fs - fb) * m.face_area,
         "depth": fs - fb, "surf": fs, "bottom": fb}
    st = OP.PlaneSWEState(P, A, m.face_mask)
    leaf = m.face_mask == 0
    pse = plane_cases.pse_eps_of(np.sqrt(m.face_area[leaf].sum() / leaf.sum()))  # mesh.appx_mesh_size()
    OP.swe_plane_init_direct_sums(st, 0.0, pse)
    # the engine's own init_direct_sums on the same raw fields
    sums_p = api.swe_plane_sums(engine, st.p["xy"], st.p["surf"], st.a["xy"], st.a["vort"], st.a["div"], st.a["area"], st.mask,
                                st.a["surf"], 0.0, pse, targets_are_sources=False)
    assert field_rel_err(sums_p["laps"], st.p["laps"]) < SUM_TOL  # the fluid starts at rest: only the PSE Laplacian is non-zero
    assert np.abs(sums_p["vel"]).max() == 0.0 and np.abs(st.p["vel"]).max() == 0.0
    gp = {k: st.p[k].copy() for k in api.PLANE_PASSIVE_FIELDS}
    ga = {k: st.a[k].copy() for k in api.PLANE_ACTIVE_FIELDS}
    OP.swe_plane_rk4_step(0.05, 0.0, 0.0, 1.0, 0.0, pse, 1, st, n_steps=3)
    api.swe_plane_rk4_step(engine, 0.05, 0.0, 0.0, 1.0, 0.0, pse, 1, gp, ga, st.mask, n_steps=3)
    errs = {"p_" + k: field_rel_err(gp[k], st.p[k]) for k in api.PLANE_PASSIVE_FIELDS}
    errs.update({"a_" + k: field_rel_err(ga[k], st.a[k]) for k in api.PLANE_ACTIVE_FIELDS})
    assert max(errs.values()) < STEP_TOL, errs
    assert np.array_equal(ga["mass"], st.a["mass"])  # mass is carried, never recomputed
