"""Gather / scatter and the GMLS surface Laplacian.

PARITY UNPINNED against Compadre (absent third-party dependency, see lpm_b200/csrc/lpmx_gmls_core.h and
oracle/gmls_oracle.py).  What is checked instead:
  CPU  the product's per-target arithmetic (compiled for the host) == the numpy restatement (independent lstsq path) to
       1e-10; convergence to the exact Laplace-Beltrami of spherical harmonics; the TC2 closed form
       (examples/sphere_swe_tc2.cpp:243-244); robustness off the unit sphere; gather/scatter restatement round trip.
  GPU  lpmx_gmls_sphere_laplacian through the C ABI == host arithmetic (1e-9: normal equations, different FMA
       contraction), identical window radii and neighbour counts; lpmx_gather/scatter_mesh_data bit-exact; SWERK2 steps
       with the built-in device provider == the oracle stepper fed by the host arithmetic."""
import numpy as np
import pytest

from conftest import field_rel_err
from gmls_util import core_host_interpolate, core_host_laplacian, harmonic_field, remesh_case
from lpm_b200 import gallery
from lpm_b200.api import PolyMesh2d
from oracle import gmls_oracle as GO


def _cloud(seed, depth):
    m = PolyMesh2d(seed, depth)
    return m, GO.gather(m.vert_xyz, m.face_xyz, m.face_mask)


@pytest.mark.parametrize("seed,depth", [("cubed", 3), ("icos", 2)])
@pytest.mark.parametrize("order", [2, 3, 4])
def test_core_arithmetic_matches_numpy_restatement(seed, depth, order):
    _, x = _cloud(seed, depth)
    f, _ = harmonic_field(x)
    p = GO.params(order)
    lap_c, eps_c, nn_c = core_host_laplacian(x, f, p)
    lap_o, eps_o, nn_o = GO.sphere_laplacian(x, f, p)
    assert np.array_equal(nn_c, nn_o)
    assert np.abs(eps_c - eps_o).max() < 1e-15
    assert field_rel_err(lap_c, lap_o) < 1e-10


def test_core_mixed_orders_and_neighbor_overrides_match_numpy():
    _, x = _cloud("cubed", 3)
    f, _ = harmonic_field(x)
    p = GO.params(4, eps_multiplier=1.8, manifold_order=2, min_neighbors=20)
    lap_c, eps_c, nn_c = core_host_laplacian(x, f, p)
    lap_o, eps_o, nn_o = GO.sphere_laplacian(x, f, p)
    assert np.array_equal(nn_c, nn_o) and field_rel_err(lap_c, lap_o) < 1e-10


def test_converges_to_exact_laplace_beltrami_of_spherical_harmonics():
    errs = {}
    for depth in (3, 4, 5):
        _, x = _cloud("cubed", depth)
        f, exact = harmonic_field(x)
        for order in (2, 4):
            lap, _, nn = core_host_laplacian(x, f, GO.params(order))
            errs[(depth, order)] = field_rel_err(lap, exact)
            assert nn.min() >= GO.params(order)["min_neighbors"]
    assert errs[(5, 2)] < 1e-2 and errs[(5, 4)] < 1e-4
    # a degree-m fit differentiates twice: O(h^(m-1)) in general (more on near-symmetric stencils); halving h, depth 4 -> 5
    assert np.log2(errs[(4, 2)] / errs[(5, 2)]) > 1.0
    assert np.log2(errs[(4, 4)] / errs[(5, 4)]) > 2.5


def test_tc2_surface_laplacian_closed_form():
    """examples/sphere_swe_tc2.cpp:243-244: lap s = (u0^2 + 2 Omega u0)(2 sin^2 - cos^2)/g ... as gallery states it"""
    tc = gallery.SphereTestCase2()
    _, x = _cloud("icos", 4)
    lap, _, _ = core_host_laplacian(x, tc.surface_exact(x), GO.params(4))
    assert field_rel_err(lap, tc.surface_laplacian_exact(x)) < 5e-4


def test_points_slightly_off_the_sphere_and_other_radii():
    """time stepping leaves |x| = 1 + O(1e-10..1e-6); the height reconstruction absorbs it.  Radius R scales lap by 1/R^2."""
    _, x = _cloud("cubed", 4)
    f, exact = harmonic_field(x)
    rng = np.random.default_rng(3)
    xo = x * (1 + 1e-7 * rng.standard_normal((x.shape[0], 1)))
    lap, _, _ = core_host_laplacian(xo, f, GO.params(4))
    assert field_rel_err(lap, exact) < 2e-3
    lap2, eps2, _ = core_host_laplacian(2.5 * x, f, GO.params(4), radius=2.5)
    lap1, eps1, _ = core_host_laplacian(x, f, GO.params(4))
    assert field_rel_err(lap2 * 2.5 ** 2, lap1) < 1e-9 and np.abs(eps2 / 2.5 - eps1).max() < 1e-14


def test_ill_conditioned_fit_is_reported_as_nan():
    """A cloud whose points all lie (to 1e-9) on one great circle: every target's neighbours are on a line in its tangent plane,
    the moment matrix has no information across it, and the fit's coefficients carry no digits.  The Cholesky pivots then span
    more than 1e12 and the core returns NaN -- on the quasi-uniform meshes the same check never fires (all finite)."""
    rng = np.random.default_rng(7)
    n = 3000
    lam = np.sort(rng.uniform(0, 2 * np.pi, n))
    xyz = np.stack([np.cos(lam), np.sin(lam), 1e-9 * rng.standard_normal(n)], axis=1)
    xyz /= np.linalg.norm(xyz, axis=1)[:, None]
    p = GO.params(4)
    lap, _, _ = core_host_laplacian(xyz, np.cos(3 * lam), p)
    assert np.isnan(lap).all()
    _, x_ok = _cloud("icos", 3)
    lap_ok, _, _ = core_host_laplacian(x_ok, harmonic_field(x_ok)[0], p)
    assert np.isfinite(lap_ok).all()


def test_gather_scatter_restatement_round_trip():
    m = PolyMesh2d("icos", 2)
    rng = np.random.default_rng(5)
    vd, fd = rng.standard_normal((m.n_verts, 3)), rng.standard_normal((m.n_faces, 3))
    g = GO.gather(vd, fd, m.face_mask)
    assert g.shape[0] == m.n_verts + m.n_face_leaves
    leaf = m.face_mask == 0
    assert np.array_equal(g[m.n_verts + m.face_leaf_idx[leaf]], fd[leaf])  # row n_verts + faces.leaf_idx(f)
    v2, f2 = GO.scatter(g, m.n_verts, m.face_mask, np.full_like(fd, -1.0))
    assert np.array_equal(v2, vd) and np.array_equal(f2[leaf], fd[leaf]) and (f2[~leaf] == -1.0).all()


def _reference_gather_scatter_scenario(gather, scatter, m):
    """tests/lpm_gather_scatter_mesh_tests.cpp:38-125: every vertex and face carries 2, the gathered copy is set to 3 and
    scattered back -> every vertex and every LEAF holds 3, every divided face still 2; the gathered points are
    n_vertices + n_leaves and no two of them coincide."""
    vert, face = np.full(m.n_verts, 2.0), np.full(m.n_faces, 2.0)
    g = gather(vert, face, m.face_mask)
    assert g.shape[0] == m.n_verts + m.n_face_leaves                       # :61
    x = gather(m.vert_xyz, m.face_xyz, m.face_mask)
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(axis=2) + np.eye(x.shape[0])
    assert (d2 > 0).all()                                                  # :63-89 n_duplicates == 0
    vert, face = scatter(np.full_like(g, 3.0), vert, face, m.face_mask)
    assert (vert == 3.0).sum() == m.n_verts                                # :103
    assert (face == 3.0).sum() == m.n_face_leaves                          # :124
    assert (face == 2.0).sum() == m.n_faces - m.n_face_leaves              # :125


@pytest.mark.parametrize("seed", ["cubed", "icos", "quad_rect"])
def test_reference_gather_scatter_scenario_on_the_restatement(seed):
    m = PolyMesh2d(seed, 2)
    _reference_gather_scatter_scenario(GO.gather, lambda g, v, f, mask: GO.scatter(g, v.shape[0], mask, f.copy()), m)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["cubed", "icos"])
def test_gpu_reference_gather_scatter_scenario(engine, seed):
    m = PolyMesh2d(seed, 2)

    def scatter(g, v, f, mask):
        vo, fo = v.copy(), f.copy()
        engine.scatter_mesh_data(g, vo, fo, mask)
        return vo, fo
    _reference_gather_scatter_scenario(engine.gather_mesh_data, scatter, m)


def test_interpolation_core_matches_numpy_and_converges():
    """Remesh interpolation (ScalarPointEvaluation at new particles): host build of the product arithmetic against the
    numpy restatement, five fields at once (two batches of kInterpFields = 4), and convergence to the exact values."""
    src, F, tgt, exact = remesh_case(3)
    for order in (1, 2, 4):
        p = GO.params(order)
        got = core_host_interpolate(src, F, tgt, p)
        assert np.abs(got - GO.sphere_interpolate(src, F, tgt, p)).max() < 1e-12, order
    e = {}
    for depth in (3, 4):
        src, F, tgt, exact = remesh_case(depth)
        for order in (2, 4):
            e[(depth, order)] = np.abs(core_host_interpolate(src, F, tgt, GO.params(order)) - exact).max()
    assert e[(4, 2)] < 5e-4 and e[(4, 4)] < 1e-5
    assert np.log2(e[(3, 2)] / e[(4, 2)]) > 2.5 and np.log2(e[(3, 4)] / e[(4, 4)]) > 4.5  # O(h^(m+1))


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("depth,order", [(3, 1), (3, 3), (4, 4), (5, 4)])
def test_gpu_interpolation_matches_host_arithmetic(engine, depth, order):
    src, F, tgt, exact = remesh_case(depth)
    p = GO.params(order)
    got = engine.gmls_sphere_interpolate(src, list(F), tgt, order)
    assert np.abs(got - core_host_interpolate(src, F, tgt, p)).max() < 1e-11
    if order == 4:
        assert np.abs(got - exact).max() < (1e-5 if depth == 4 else 5e-7)
    if depth == 3:
        assert np.abs(got - GO.sphere_interpolate(src, F, tgt, p)).max() < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("seed,depth,order", [("cubed", 3, 2), ("icos", 3, 3), ("cubed", 4, 4), ("icos", 5, 4)])
def test_gpu_laplacian_matches_host_arithmetic_and_numpy(engine, seed, depth, order):
    _, x = _cloud(seed, depth)
    f, exact = harmonic_field(x)
    p = GO.params(order)
    lap, eps, nn = engine.gmls_sphere_laplacian(x, f, order, diagnostics=True)
    lap_c, eps_c, nn_c = core_host_laplacian(x, f, p)
    assert np.array_equal(nn, nn_c)
    assert np.abs(eps - eps_c).max() < 1e-14
    assert field_rel_err(lap, lap_c) < 1e-9
    if depth <= 3:
        lap_o, _, nn_o = GO.sphere_laplacian(x, f, p)
        assert np.array_equal(nn, nn_o) and field_rel_err(lap, lap_o) < 1e-9
    if order == 4 and depth >= 4:
        assert field_rel_err(lap, exact) < 1e-3


@pytest.mark.gpu
def test_gpu_laplacian_layout_left_device_pointers_and_full_size(engine):
    import torch
    from lpm_b200.api import LAYOUT_LEFT
    _, x = _cloud("cubed", 6)  # 49k points
    f, exact = harmonic_field(x)
    dev = torch.device("cuda", 0)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
    lap = engine.gmls_sphere_laplacian(xt, torch.from_numpy(f).to(dev), 4, layout=LAYOUT_LEFT)
    engine.sync()
    assert field_rel_err(lap.cpu().numpy(), exact) < 2e-5
    # a shuffled copy of the cloud gives the same values at the same points (the grid sort is order independent
    # up to the summation order inside a cell)
    perm = np.random.default_rng(9).permutation(x.shape[0])
    lap_p = engine.gmls_sphere_laplacian(x[perm], f[perm], 4)
    assert field_rel_err(lap_p, lap.cpu().numpy()[perm]) < 1e-9


@pytest.mark.gpu
def test_gpu_params_are_validated(engine):
    from lpm_b200.api import LpmxError, gmls_params
    _, x = _cloud("cubed", 2)
    f, _ = harmonic_field(x)
    for bad in (gmls_params(5), gmls_params(3, manifold_weight_pwr=3.0), gmls_params(3, min_neighbors=64),
                gmls_params(3, topo_dim=3)):
        with pytest.raises(LpmxError):
            engine.gmls_sphere_laplacian(x, f, bad)
    # too few neighbours for the polynomial space: rank-deficient systems give NaN, not garbage
    lap = engine.gmls_sphere_laplacian(x, f, gmls_params(4, min_neighbors=4, eps_multiplier=1.0))
    assert np.isnan(lap).any()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,depth", [("cubed", 3), ("icos", 3)])
def test_gpu_gather_scatter_bit_exact(engine, seed, depth):
    m = PolyMesh2d(seed, depth)
    rng = np.random.default_rng(11)
    for shape in ((), (3,)):
        vd, fd = rng.standard_normal((m.n_verts,) + shape), rng.standard_normal((m.n_faces,) + shape)
        g = engine.gather_mesh_data(vd, fd, m.face_mask)
        assert np.array_equal(g, GO.gather(vd, fd, m.face_mask))
        vo, fo = np.zeros_like(vd), np.full_like(fd, -1.0)
        engine.scatter_mesh_data(g, vo, fo, m.face_mask)
        rv, rf = GO.scatter(g, m.n_verts, m.face_mask, np.full_like(fd, -1.0))
        assert np.array_equal(vo, rv) and np.array_equal(fo, rf)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,depth", [("cubed", 3), ("icos", 2)])
def test_gpu_swe_rk2_with_builtin_provider_matches_oracle_with_host_gmls(engine, oracle, seed, depth):
    """Two SWERK2 steps: engine with lpmx_gmls_swe_laplacian (gather -> GMLS -> scatter on the device) against the
    oracle stepper whose Laplacian callback runs the same arithmetic on the host."""
    from lpm_b200.api import ACTIVE_FIELDS, PASSIVE_FIELDS, swe_rk2_step
    from test_gpu_parity_swe_rk2 import G, OMEGA, compare, tc2_state
    m = PolyMesh2d(seed, depth)
    p = GO.params(3)

    def host_gmls(stage, px, psurf, ax, asurf, amask):
        x = GO.gather(px, ax, amask)
        s = GO.gather(psurf, asurf, amask)
        lap, _, _ = core_host_laplacian(x, s, p)
        return GO.scatter(lap, px.shape[0], amask, np.zeros(ax.shape[0]))

    ref = tc2_state(oracle, m, eps=0.05, div_amp=0.1)
    # the reference enters the first step with the Laplacian of the initial state (SWERK2 constructor, rk2_impl.hpp:56-77)
    ref.p["laps"], ref.a["laps"] = host_gmls(0, ref.p["xyz"], ref.p["surf"], ref.a["xyz"], ref.a["surf"], m.face_mask)
    got_p = {k: ref.p[k].copy() for k in PASSIVE_FIELDS}
    got_a = {k: ref.a[k].copy() for k in ACTIVE_FIELDS}
    oracle.swe_rk2_step(0.01, OMEGA, G, 0.05, ref, laps_fn=host_gmls, n_steps=2)
    swe_rk2_step(engine, 0.01, OMEGA, G, 0.05, got_p, got_a, m.face_mask, laplacian=engine.gmls_provider(3), n_steps=2)
    leaf = m.face_mask == 0
    assert field_rel_err(got_p["laps"], ref.p["laps"]) < 1e-8
    assert field_rel_err(got_a["laps"][leaf], ref.a["laps"][leaf]) < 1e-8
    for k in ("xyz", "vort", "div", "depth"):
        assert field_rel_err(got_p[k], ref.p[k]) < 1e-9, k
        assert field_rel_err(got_a[k][leaf], ref.a[k][leaf]) < 1e-9, k
