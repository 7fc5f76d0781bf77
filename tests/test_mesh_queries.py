"""Mesh queries of the adaptive tree (lpmx_mesh_leaf_edges_from_parent / ccw_edges_around_face / ccw_adjacent_faces /
neighbors_flag / locate; host code like the mesh, src/mesh/lpm_polymesh2d.hpp:262-552).  The first tests are the REFERENCE'S OWN
known answers (tests/lpm_polymesh2d_function_tests.cpp:50-252: QuadRectSeed depth 0, face 0 divided, then its first kid), which
also pin the planar quad divider and the edge tree against values the reference asserts."""
import os
import subprocess

import numpy as np
import pytest

from lpm_b200.api import LpmxError, PolyMesh2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _divide(m, faces):
    f = np.zeros(m.n_faces, dtype=np.uint8)
    f[list(faces)] = 1
    n, outcome = m.divide_flagged_faces(f)
    assert n == len(faces) and outcome == PolyMesh2d.AMR_DIVIDED_ALL


@pytest.fixture()
def qr0():
    m = PolyMesh2d("quad_rect", 0, 1.0, amr_buffer=3, amr_limit=3)
    assert m.n_faces == 4                      # :57
    _divide(m, [0])
    assert m.face_kids[0][0] == 4              # :64
    _divide(m, [int(m.face_kids[0][0])])
    return m


def test_reference_edge_tree_after_two_divisions(qr0):
    m = qr0                                    # :91-100
    assert m.n_faces == 12
    assert list(m.edge_kids[0]) == [12, 13] and list(m.edge_kids[12]) == [24, 25]
    assert m.edge_lefts[24] == 8 and m.edge_rights[24] == -1 and m.edge_lefts[25] == 9
    assert np.allclose(m.face_xyz[8], [-7.0 / 8, 7.0 / 8], rtol=0, atol=1e-15)


def test_reference_leaf_edges_and_adjacency(qr0):
    m = qr0                                    # :213-238
    assert list(m.get_leaf_edges_from_parent(0)) == [24, 25, 13]
    assert list(m.ccw_edges_around_face(7)) == [29, 28, 21, 17, 18]
    assert list(m.ccw_adjacent_faces(5)) == [-1, 1, 6, 10, 9]


def test_reference_point_location(qr0):
    m = qr0                                    # :163-191, :240-246
    assert list(m.locate_face_containing_pt(m.face_xyz)) == [10, 1, 2, 3, 8, 5, 6, 7, 8, 9, 10, 11]
    assert list(m.locate_face_containing_pt(m.vert_xyz)) == [8, 5, 1, 1, 2, 2, 3, 7, 6, 9, 5, 6, 11, 10, 8, 9, 10, 8, 8]
    qp = np.array([[-0.875, 0.875]])
    assert m.locate_pt_walk_search(qp, 2)[0] == 8
    assert m.nearest_root_face(qp)[0] == 0
    assert m.locate_pt_tree_search(qp, 0)[0] == 8
    assert m.locate_face_containing_pt(qp)[0] == 8
    assert m.locate_face_containing_pt(np.array([[1.5, 0.2], [0.3, -1.01]])).tolist() == [-1, -1]   # pt_is_outside_mesh
    with pytest.raises(LpmxError):
        m.locate_pt_walk_search(qp, 0)         # leaf-only (:381-382)


@pytest.mark.parametrize("seed,depth,nfv", [("icos", 3, 3), ("cubed", 3, 4), ("quad_rect", 2, 4), ("tri_hex", 2, 3)])
def test_uniform_mesh_adjacency_is_symmetric_and_leaves_find_themselves(seed, depth, nfv):
    m = PolyMesh2d(seed, depth)
    leaves = np.nonzero(m.face_mask == 0)[0]
    adj = {int(f): [int(a) for a in m.ccw_adjacent_faces(int(f))] for f in leaves}
    for f, a in adj.items():
        assert len(a) == nfv
        for b in a:
            if b >= 0:
                assert m.face_mask[b] == 0 and f in adj[b]
            else:
                assert m.ndim == 2           # only planar meshes have a boundary
        assert list(m.ccw_edges_around_face(f)) == list(m.face_edges[f])   # nothing divided: the face's own edges
    assert np.array_equal(m.locate_face_containing_pt(m.face_xyz[leaves]), leaves)
    # a divided face resolves to one of its descendants
    roots = np.nonzero(m.face_parent < 0)[0]
    found = m.locate_face_containing_pt(m.face_xyz[roots])
    for r, f in zip(roots, found):
        p = int(f)
        while m.face_parent[p] >= 0:
            p = int(m.face_parent[p])
        assert p == r and m.face_mask[f] == 0


def _balanced(m):
    """no leaf has a leaf neighbour more than one level finer"""
    for f in np.nonzero(m.face_mask == 0)[0]:
        for a in m.ccw_adjacent_faces(int(f)):
            if a >= 0 and m.face_level[a] > m.face_level[f] + 1:
                return False
    return True


@pytest.mark.parametrize("seed", ["icos", "cubed"])
def test_neighbors_flag_restores_two_to_one_balance(seed):
    m = PolyMesh2d(seed, 2, amr_buffer=3, amr_limit=3)
    target = int(np.nonzero(m.face_mask == 0)[0][5])
    for _ in range(3):                               # refine towards one corner three times: unbalanced
        _divide(m, [target])
        target = int(m.face_kids[target][0])
    assert not _balanced(m)
    for _ in range(6):
        flags = np.zeros(m.n_faces, dtype=np.uint8)
        m.neighbors_flag(flags)
        # the functor as coded also flags divided faces (it never looks at the mask); dividing is for leaves only
        todo = [int(i) for i in np.nonzero(flags)[0] if m.face_mask[i] == 0]
        if not todo:
            break
        _divide(m, todo)
    assert _balanced(m)
    assert abs(m.face_area[m.face_mask == 0].sum() - 4 * np.pi) < 1e-12


def test_neighbors_flag_only_adds_and_counts_new_flags(qr0):
    flags = np.zeros(qr0.n_faces, dtype=np.uint8)
    flags[3] = 1
    added = qr0.neighbors_flag(flags)
    assert flags[3] == 1 and added == int(flags.sum()) - 1
    assert np.nonzero(flags)[0].tolist() == [0, 3]    # the divided root face 0 sees its level-3 descendants across its own edges
    assert qr0.neighbors_flag(flags) == 0


def test_cpp_shim_runs_the_reference_function_test():
    from lpm_b200 import build
    build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "mesh_queries_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "mesh_queries_check.cpp"), "-L" + os.path.join(ROOT, "lpm_b200"), "-llpmx",
                    "-Wl,-rpath," + os.path.join(ROOT, "lpm_b200")], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout + p.stderr
