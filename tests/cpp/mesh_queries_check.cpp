// mesh_queries_check.cpp -- the assertions of the reference's own mesh-function test
// (tests/lpm_polymesh2d_function_tests.cpp:50-252: a QuadRectSeed mesh of depth 0 whose face 0 and then that face's first kid
// are divided) run against the C++ shim: edge-tree layout, get_leaf_edges_from_parent, ccw_edges_around_face,
// ccw_adjacent_faces, the four point-location functions, plus NeighborsFlag through Refinement::iterate.
// Host-only (the mesh is host code); prints "ok" or the first failed check.
#include <cmath>
#include <cstdio>
#include <vector>

#include "lpm/lpm.hpp"

using namespace Lpm;

#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      std::printf("FAILED line %d: %s\n", __LINE__, #cond);          \
      return 1;                                                      \
    }                                                                \
  } while (0)

int main() {
  try {
    PolyMeshParameters<QuadRectSeed> params(0, 1.0, 3, 3);
    PolyMesh2d<QuadRectSeed> qr0(params);
    Logger logger("mesh_queries_check", Log::none);
    CHECK(qr0.n_faces_host() == 4);
    mask_view_type flags("flags", qr0.faces.area.extent(0));
    auto divide = [&](Index f) {
      for (Index i = 0; i < (Index)flags.extent(0); ++i) flags(i) = 0;
      flags(f) = 1;
      qr0.divide_flagged_faces(flags, logger);
    };
    divide(0);
    CHECK(qr0.faces.kids(0, 0) == 4);
    divide(qr0.faces.kids(0, 0));
    // :91-100
    CHECK(qr0.n_faces_host() == 12);
    CHECK(qr0.edges.kids(0, 0) == 12 && qr0.edges.kids(0, 1) == 13);
    CHECK(qr0.edges.kids(12, 0) == 24 && qr0.edges.kids(12, 1) == 25);
    CHECK(qr0.edges.lefts(24) == 8 && qr0.edges.rights(24) == LPM_NULL_IDX && qr0.edges.lefts(25) == 9);
    CHECK(std::fabs(qr0.faces.phys_crds.view(8, 0) + 7.0 / 8) < 1e-15 && std::fabs(qr0.faces.phys_crds.view(8, 1) - 7.0 / 8) < 1e-15);
    // :213-238
    Index list[64];
    Int n = 0;
    qr0.get_leaf_edges_from_parent(list, n, 0);
    CHECK(n == 3 && list[0] == 24 && list[1] == 25 && list[2] == 13);
    qr0.ccw_edges_around_face(list, n, 7);
    CHECK(n == 5 && list[0] == 29 && list[1] == 28 && list[2] == 21 && list[3] == 17 && list[4] == 18);
    qr0.ccw_adjacent_faces(list, n, 5);
    CHECK(n == 5 && list[0] == LPM_NULL_IDX && list[1] == 1 && list[2] == 6 && list[3] == 10 && list[4] == 9);
    // :240-246
    const Real qp[2] = {-0.875, 0.875};
    CHECK(qr0.locate_pt_walk_search(qp, 2) == 8);
    CHECK(qr0.nearest_root_face(qp) == 0);
    CHECK(qr0.locate_pt_tree_search(qp, 0) == 8);
    CHECK(qr0.locate_face_containing_pt(qp) == 8);
    // :163-191
    const Index face_correct[12] = {10, 1, 2, 3, 8, 5, 6, 7, 8, 9, 10, 11};
    for (Index i = 0; i < 12; ++i) {
      const Real p[2] = {qr0.faces.phys_crds.view(i, 0), qr0.faces.phys_crds.view(i, 1)};
      CHECK(qr0.locate_face_containing_pt(p) == face_correct[i]);
    }
    const Index vert_correct[19] = {8, 5, 1, 1, 2, 2, 3, 7, 6, 9, 5, 6, 11, 10, 8, 9, 10, 8, 8};
    CHECK(qr0.n_vertices_host() == 19);
    for (Index i = 0; i < 19; ++i) {
      const Real p[2] = {qr0.vertices.phys_crds.view(i, 0), qr0.vertices.phys_crds.view(i, 1)};
      CHECK(qr0.locate_face_containing_pt(p) == vert_correct[i]);
    }
    const Real outside[2] = {1.5, 0.2};
    CHECK(qr0.locate_face_containing_pt(outside) == LPM_NULL_IDX);
    // NeighborsFlag as coded: the level-3 faces 8..11 sit in the outer corner and touch only level-2 faces and the boundary,
    // so no LEAF is out of balance; the functor does not look at the mask, though, and the divided root face 0 (level 1) sees
    // the level-3 leaves across its own leaf edges -> it is the one face flagged
    Refinement<QuadRectSeed> refine(qr0);
    NeighborsFlag<QuadRectSeed> nf(refine.flags, qr0);
    refine.iterate(0, qr0.n_faces_host(), nf);
    std::vector<Index> on;
    for (Index i = 0; i < qr0.n_faces_host(); ++i)
      if (refine.flags(i)) on.push_back(i);
    std::printf("flagged:");
    for (Index i : on) std::printf(" %d", (int)i);
    std::printf(" count %d\n", (int)refine.count[0]);
    CHECK((Index)on.size() == refine.count[0]);
    CHECK(on.size() == 1 && on[0] == 0);
    std::printf("ok\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 4;
  }
}
