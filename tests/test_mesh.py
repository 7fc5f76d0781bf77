"""CPU: the host mesh generator (input generator of every config) against
  * golden fixtures produced by the REFERENCE's own mesh classes compiled in place (PolyMesh2d<Seed>::tree_init of
    /root/reference/src/mesh -> oracle/_ref/liblpm_ref_mesh.so, tests/golden/make_mesh_golden.py): EVERY array bit-exact;
  * the compiled reference live, at depths beyond the fixtures (where oracle/_ref and /root/reference are present);
  * an independent pure-Python replay of the dividers (oracle/mesh_oracle.py), a second witness;
  * the reference's own mesh tests: allocation counts (tests/lpm_polymesh_tests.cpp:76-123), total area 4 pi;
  * structural invariants of the quad-tree."""
import os

import numpy as np
import pytest

from lpm_b200.api import PolyMesh2d, max_allocations

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("icos", 0), ("icos", 1), ("icos", 2), ("icos", 3), ("cubed", 0), ("cubed", 1), ("cubed", 2), ("cubed", 3),
         ("cubed", 4)]
INT_ARRAYS = ["edge_origs", "edge_dests", "edge_lefts", "edge_rights", "edge_parents", "edge_kids", "face_verts",
              "face_edges", "face_parent", "face_kids", "face_level", "face_leaf_idx", "face_mask"]
REAL_ARRAYS = ["vert_xyz", "vert_lag_xyz", "face_xyz", "face_lag_xyz", "face_area"]


@pytest.mark.parametrize("seed,depth", CASES)
def test_mesh_matches_golden_bit_exact(seed, depth):
    g = np.load(os.path.join(GOLDEN, f"mesh_{seed}_{depth}.npz"))
    m = PolyMesh2d(seed, depth)
    for k in INT_ARRAYS:
        assert np.array_equal(g[k], getattr(m, k)), k
    for k in REAL_ARRAYS:  # same libm, no FMA contraction on either side: bitwise equality
        assert np.array_equal(g[k].view(np.int64), getattr(m, k).view(np.int64)), k


def test_embedded_seed_tables_equal_reference_dat_files():
    t = np.load(os.path.join(GOLDEN, "seed_tables.npz"))
    for seed, nv in (("icos", 12), ("cubed", 8)):
        m = PolyMesh2d(seed, 0)
        assert np.array_equal(t[f"{seed}_crds"][:nv], m.vert_xyz)
        assert np.array_equal(t[f"{seed}_crds"][nv:], m.face_xyz)
        assert np.array_equal(t[f"{seed}_edges"][:, 0], m.edge_origs)
        assert np.array_equal(t[f"{seed}_edges"][:, 1], m.edge_dests)
        assert np.array_equal(t[f"{seed}_edges"][:, 2], m.edge_lefts)
        assert np.array_equal(t[f"{seed}_edges"][:, 3], m.edge_rights)
        assert np.array_equal(t[f"{seed}_face_verts"], m.face_verts)
        assert np.array_equal(t[f"{seed}_face_edges"], m.face_edges)


@pytest.mark.skipif(not os.path.isdir("/root/reference/mesh_seeds"), reason="reference not mounted")
def test_live_python_restatement_from_reference_seed_files():
    from oracle import mesh_oracle
    for seed, depth in (("icos", 2), ("cubed", 3), ("quad_rect", 3), ("tri_hex", 2)):
        ref = mesh_oracle.TreeMesh(seed, depth).arrays()
        m = PolyMesh2d(seed, depth)
        for k, v in ref.items():
            assert np.array_equal(v, getattr(m, k)), k


def _ref_mesh_or_skip():
    from oracle import ref_mesh
    if not ref_mesh.available():
        pytest.skip("oracle/_ref/liblpm_ref_mesh.so or /root/reference not present (build container only)")
    return ref_mesh


@pytest.mark.parametrize("seed,depth,radius", [("icos", 4, 1.0), ("icos", 5, 1.0), ("cubed", 5, 1.0), ("cubed", 6, 1.0),
                                               ("quad_rect", 5, 3.0), ("tri_hex", 4, 2.5)])
def test_live_reference_mesh_classes_every_array_bit_exact(seed, depth, radius):
    """The reference's PolyMesh2d<Seed>(PolyMeshParameters(depth, radius)) (TriFace / QuadFace dividers of
    src/mesh/lpm_faces_impl.hpp:284-574, Edges::divide of lpm_edges.cpp:58-125, tree_init of lpm_polymesh2d_impl.hpp:25-42),
    compiled in place, against the product's generator: insertion order, every index array, every coordinate and area."""
    ref_mesh = _ref_mesh_or_skip()
    r = ref_mesh.RefMesh(seed, depth, radius)
    a = r.arrays()
    c = r.counts()
    r.close()
    m = PolyMesh2d(seed, depth, radius=radius)
    assert (m.n_verts, m.n_edges, m.n_faces, m.n_face_leaves) == (c["n_verts"], c["n_edges"], c["n_faces"], c["n_leaves"])
    assert (c["nmaxverts"], c["nmaxedges"], c["nmaxfaces"]) == max_allocations(seed, depth)
    for k in INT_ARRAYS:
        assert np.array_equal(a[k], getattr(m, k)), k
    for k in REAL_ARRAYS:
        assert np.array_equal(a[k].view(np.int64), getattr(m, k).view(np.int64)), k
    # Faces::crd_inds is the identity on a freshly built tree (face i owns coordinate row i)
    assert np.array_equal(a["face_crd_idx"], np.arange(m.n_faces, dtype=np.int32))


def test_golden_fixtures_are_what_the_compiled_reference_produces():
    """The committed fixtures against a fresh run of the compiled reference (guards the fixtures themselves)."""
    ref_mesh = _ref_mesh_or_skip()
    for seed, depth in CASES:
        g = np.load(os.path.join(GOLDEN, f"mesh_{seed}_{depth}.npz"))
        r = ref_mesh.RefMesh(seed, depth)
        a = r.arrays()
        r.close()
        for k in g.files:
            assert np.array_equal(g[k], a[k]), (seed, depth, k)


@pytest.mark.parametrize("seed,depth", [("icos", 3), ("cubed", 3), ("icos", 5), ("cubed", 6)])
def test_counts_equal_max_allocations_and_area_is_4pi(seed, depth):
    """tests/lpm_polymesh_tests.cpp:76-123: nh == nmax for all three element kinds; area = 4 pi."""
    m = PolyMesh2d(seed, depth)
    assert (m.n_verts, m.n_edges, m.n_faces) == max_allocations(seed, depth)
    area = 0.0
    for a in m.face_area:  # sequential, as surface_area_host does
        area += a
    assert abs(area - 4 * np.pi) < 31e-14
    assert m.n_face_leaves == (20 if seed == "icos" else 6) * 4 ** depth


@pytest.mark.parametrize("seed,depth", [("icos", 4), ("cubed", 5)])
def test_tree_invariants(seed, depth):
    m = PolyMesh2d(seed, depth)
    leaf = m.face_mask == 0
    # mask <=> divided <=> area 0; leaves have positive area and level depth+1 (root level 1)
    assert np.array_equal(~leaf, m.face_kids[:, 0] > 0)
    assert (m.face_area[~leaf] == 0).all() and (m.face_area[leaf] > 0).all()
    assert (m.face_level[leaf] == depth + 1).all()
    # leaf_idx is the exclusive scan of leaf flags
    assert np.array_equal(m.face_leaf_idx, np.concatenate([[0], np.cumsum(leaf)[:-1]]).astype(np.int32))
    # kids point back to their parent
    for k in range(4):
        kids = m.face_kids[~leaf, k]
        assert np.array_equal(m.face_parent[kids], np.nonzero(~leaf)[0])
    # every leaf edge separates two leaf faces that list it
    eleaf = m.edge_kids[:, 0] <= 0
    for side in (m.edge_lefts, m.edge_rights):
        f = side[eleaf]
        assert leaf[f].all()
        assert (m.face_edges[f] == np.nonzero(eleaf)[0][:, None]).any(axis=1).all()
    # Euler characteristic of the leaf mesh: V - E + F = 2
    assert m.n_verts - eleaf.sum() + leaf.sum() == 2
    # all particles on the unit sphere
    assert np.abs(np.linalg.norm(m.vert_xyz, axis=1) - 1).max() < 4e-16
    assert np.abs(np.linalg.norm(m.face_xyz, axis=1) - 1).max() < 4e-16
    # face vertices are counter-clockwise seen from outside
    v = m.vert_xyz[m.face_verts[leaf]]
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 1])
    assert ((n * m.face_xyz[leaf]).sum(axis=1) > 0).all()


# ---- planar seeds (QuadRectSeed, TriHexSeed; PlaneGeometry) -----------------------------------------------------------------
PLANE_CASES = [("quad_rect", 0, 1.0), ("quad_rect", 2, 1.0), ("quad_rect", 3, 4.0), ("quad_rect", 4, 6.0),
               ("tri_hex", 0, 1.0), ("tri_hex", 2, 1.0), ("tri_hex", 3, 6.0)]


@pytest.mark.parametrize("seed,depth,radius", PLANE_CASES)
def test_planar_mesh_matches_golden_bit_exact(seed, depth, radius):
    g = np.load(os.path.join(GOLDEN, f"mesh_{seed}_{depth}_r{radius:g}.npz"))
    m = PolyMesh2d(seed, depth, radius=radius)
    assert m.ndim == 2 and m.vert_xyz.shape[1] == 2 and m.face_xyz.shape[1] == 2
    for k in INT_ARRAYS:
        assert np.array_equal(g[k], getattr(m, k)), k
    for k in REAL_ARRAYS:
        assert np.array_equal(g[k].view(np.int64), getattr(m, k).view(np.int64)), k


def test_embedded_planar_seed_tables_equal_reference_dat_files():
    t = np.load(os.path.join(GOLDEN, "seed_tables.npz"))
    for seed, nv in (("quad_rect", 9), ("tri_hex", 7)):
        m = PolyMesh2d(seed, 0)
        assert np.array_equal(t[f"{seed}_crds"][:nv], m.vert_xyz)
        assert np.array_equal(t[f"{seed}_crds"][nv:], m.face_xyz)
        assert np.array_equal(t[f"{seed}_edges"], np.stack([m.edge_origs, m.edge_dests, m.edge_lefts, m.edge_rights], axis=1))
        assert np.array_equal(t[f"{seed}_face_verts"], m.face_verts)
        assert np.array_equal(t[f"{seed}_face_edges"], m.face_edges)


@pytest.mark.parametrize("seed,depth,radius,area", [("tri_hex", 3, 1.0, 1.5 * np.sqrt(3.0)), ("quad_rect", 3, 4.0, 64.0),
                                                      ("quad_rect", 5, 6.0, 144.0)])
def test_planar_counts_equal_max_allocations_and_area(seed, depth, radius, area):
    """tests/lpm_polymesh_tests.cpp:30-67: nh == nmax for all three element kinds; the hexagon has area 3 sqrt(3)/2 r^2,
    the square (2 r)^2 (FloatingPoint::equiv, i.e. within zero_tol scaled -- here 1e-13 relative)."""
    m = PolyMesh2d(seed, depth, radius=radius)
    assert (m.n_verts, m.n_edges, m.n_faces) == max_allocations(seed, depth)
    total = 0.0
    for a in m.face_area:
        total += a
    assert abs(total - area) < 1e-13 * area
    assert m.n_face_leaves == (4 if seed == "quad_rect" else 6) * 4 ** depth


@pytest.mark.parametrize("seed", ["quad_rect", "tri_hex"])
def test_planar_tree_invariants(seed):
    depth, radius = 4, 2.0
    m = PolyMesh2d(seed, depth, radius=radius)
    leaf = m.face_mask == 0
    assert np.array_equal(~leaf, m.face_kids[:, 0] > 0)
    assert (m.face_area[~leaf] == 0).all() and (m.face_area[leaf] > 0).all()
    assert (m.face_level[leaf] == depth + 1).all()
    # free boundary: boundary edges have no right face; interior leaf edges separate two leaves
    eleaf = m.edge_kids[:, 0] <= 0
    boundary = m.edge_rights < 0
    assert leaf[m.edge_lefts[eleaf]].all() and leaf[m.edge_rights[eleaf & ~boundary]].all()
    n_b = (eleaf & boundary).sum()
    assert n_b == (8 if seed == "quad_rect" else 6) * 2 ** depth
    # Euler characteristic of a disc: V - E + F = 1
    assert m.n_verts - eleaf.sum() + leaf.sum() == 1
    # boundary vertices sit on the square / inside the circumscribed circle; faces are counter-clockwise
    if seed == "quad_rect":
        assert np.abs(m.vert_xyz).max() == radius
    else:
        assert np.linalg.norm(m.vert_xyz, axis=1).max() <= radius * (1 + 1e-15)
    v = m.vert_xyz[m.face_verts[leaf]]
    cr = (v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 1, 1]) - (v[:, 1, 1] - v[:, 0, 1]) * (v[:, 2, 0] - v[:, 1, 0])
    assert (cr > 0).all()
    # Lagrangian == physical at build time
    assert np.array_equal(m.vert_xyz, m.vert_lag_xyz) and np.array_equal(m.face_xyz, m.face_lag_xyz)


_REF_GEO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "liblpm_ref_geometry.so")


@pytest.mark.skipif(not os.path.exists(_REF_GEO), reason="oracle/_ref/liblpm_ref_geometry.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,depth", [("cubed", 0), ("icos", 0), ("cubed", 3), ("icos", 3), ("quad_rect", 2), ("tri_hex", 2)])
def test_generator_geometry_equals_reference_functions_compiled_in_place(seed, depth):
    """The reference's own Geo::polygon_area / barycenter / midpoint (src/lpm_geometry.hpp, compiled in place without FMA
    contraction: oracle/ref_geometry_driver.cpp) evaluated on the generator's vertices: every leaf area, every face centre
    created by a division and every edge midpoint equal the generator's values BIT FOR BIT.  (The seed faces' centres are
    read from the seed file, not computed.)"""
    import ctypes
    L = ctypes.CDLL(_REF_GEO)
    d, dp = ctypes.c_double, ctypes.POINTER(ctypes.c_double)
    L.ref_sphere_polygon_area.restype = d
    L.ref_plane_polygon_area.restype = d

    def P(a):
        return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    m = PolyMesh2d(seed, depth)
    sph = m.ndim == 3
    area_fn = L.ref_sphere_polygon_area if sph else L.ref_plane_polygon_area
    ctr_fn = L.ref_sphere_barycenter if sph else L.ref_plane_barycenter
    mid_fn = L.ref_sphere_midpoint if sph else L.ref_plane_midpoint
    n_checked = 0
    for f in np.nonzero(m.face_mask == 0)[0]:
        vs = np.ascontiguousarray(m.vert_xyz[m.face_verts[f]])
        ctr = np.ascontiguousarray(m.face_xyz[f])
        assert area_fn(P(ctr), P(vs), len(vs)) == m.face_area[f]
        if m.face_parent[f] >= 0:
            out = np.zeros(m.ndim)
            ctr_fn(P(out), P(vs), len(vs))
            assert np.array_equal(out, ctr)
            n_checked += 1
    for e in np.nonzero(m.edge_kids[:, 0] > 0)[0]:
        out = np.zeros(m.ndim)
        mid_fn(P(out), P(m.vert_xyz[m.edge_origs[e]]), P(m.vert_xyz[m.edge_dests[e]]))
        assert np.array_equal(out, m.vert_xyz[m.edge_dests[m.edge_kids[e][0]]])
        n_checked += 1
    assert n_checked > 0 or depth == 0


def test_compiled_reference_reads_the_rewritten_seed_files_identically():
    """Where /root/reference is not mounted (the GPU box) oracle/ref_mesh.py rewrites the four mesh_seeds/*.dat files from
    tests/golden/seed_tables.npz for the reference's own parser (MeshSeed<Seed>::read_file).  In a fresh process that is forced
    onto the rewritten files, the compiled reference must build the committed fixtures bit for bit."""
    import subprocess
    import sys
    ref_mesh = _ref_mesh_or_skip()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import numpy as np, os, sys\n"
            "sys.path.insert(0, %r)\n"
            "from oracle import ref_mesh\n"
            "for seed, depth, r in (('icos', 2, 1.0), ('cubed', 3, 1.0), ('quad_rect', 2, 1.0), ('tri_hex', 2, 1.0)):\n"
            "    f = 'mesh_%%s_%%d%%s.npz' %% (seed, depth, '' if seed in ('icos', 'cubed') else '_r1')\n"
            "    g = np.load(os.path.join(%r, f))\n"
            "    m = ref_mesh.RefMesh(seed, depth, r)\n"
            "    a = m.arrays()\n"
            "    assert all(np.array_equal(g[k], a[k]) for k in g.files), (seed, depth)\n"
            "assert os.environ['LPM_ORACLE_SEED_DIR'] != ref_mesh.SEED_DIR\n"
            "print('ok')\n") % (root, GOLDEN)
    env = dict(os.environ, LPM_ORACLE_FORCE_REWRITTEN_SEEDS="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout + p.stderr
