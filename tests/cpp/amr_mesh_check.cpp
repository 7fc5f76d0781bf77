// amr_mesh_check.cpp -- host-only check of the C++ shim's PolyMesh2d::divide_flagged_faces (no engine, no GPU):
// builds a mesh with an AMR buffer, moves nothing, runs `passes` refinement passes that flag every third leaf, and prints
// the counts and an FNV-1a hash of every mesh array in use.  tests/test_amr.py does the same through the ctypes binding
// and compares.  Usage: amr_mesh_check <icos|cubed> <depth> <amr> <passes>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "lpm/lpm.hpp"

using namespace Lpm;

static uint64_t fnv(const void* p, size_t bytes, uint64_t h = 1469598103934665603ULL) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ULL;
  return h;
}

template <typename Seed>
int run(int depth, int amr, int passes) {
  PolyMeshParameters<Seed> params(depth, Seed::geo::ndim == 2 ? 3.0 : 1.0, amr, amr);  // planar meshes: radius 3
  PolyMesh2d<Seed> mesh(params);
  Logger logger("amr_mesh_check", Log::none);
  mask_view_type flags("flags", mesh.faces.area.extent(0));
  const scalar_view_type area_alias = mesh.faces.area;  // a shallow copy taken BEFORE refinement must stay valid
  for (int p = 0; p < passes; ++p) {
    int leaf = 0;
    for (Index i = 0; i < (Index)flags.extent(0); ++i) flags(i) = 0;
    for (Index i = 0; i < mesh.n_faces_host(); ++i)
      if (!mesh.faces.mask(i)) flags(i) = (leaf++ % 3 == 0) ? 1 : 0;
    mesh.divide_flagged_faces(flags, logger);
  }
  if (area_alias.data() != mesh.faces.area.data()) return 3;
  const Index nv = mesh.n_vertices_host(), ne = mesh.n_edges_host(), nf = mesh.n_faces_host();
  std::printf("%d %d %d %d %d %d\n", nv, ne, nf, mesh.faces.n_leaves_host(), mesh.edges.n_leaves_host(), logger.count(Log::warn));
  constexpr int nd = Seed::geo::ndim;
  uint64_t h = fnv(mesh.vertices.phys_crds.view.data(), sizeof(Real) * nd * nv);
  h = fnv(mesh.vertices.lag_crds.view.data(), sizeof(Real) * nd * nv, h);
  h = fnv(mesh.faces.phys_crds.view.data(), sizeof(Real) * nd * nf, h);
  h = fnv(mesh.faces.lag_crds.view.data(), sizeof(Real) * nd * nf, h);
  h = fnv(mesh.faces.area.data(), sizeof(Real) * nf, h);
  h = fnv(mesh.faces.mask.data(), nf, h);
  h = fnv(mesh.faces.verts.data(), sizeof(Index) * Seed::faceKind::nverts * nf, h);
  h = fnv(mesh.faces.edges.data(), sizeof(Index) * Seed::faceKind::nverts * nf, h);
  h = fnv(mesh.faces.kids.data(), sizeof(Index) * 4 * nf, h);
  h = fnv(mesh.faces.parent.data(), sizeof(Index) * nf, h);
  h = fnv(mesh.faces.level.data(), sizeof(Index) * nf, h);
  h = fnv(mesh.faces.leaf_idx.data(), sizeof(Index) * nf, h);
  h = fnv(mesh.edges.origs.data(), sizeof(Index) * ne, h);
  h = fnv(mesh.edges.dests.data(), sizeof(Index) * ne, h);
  h = fnv(mesh.edges.lefts.data(), sizeof(Index) * ne, h);
  h = fnv(mesh.edges.rights.data(), sizeof(Index) * ne, h);
  h = fnv(mesh.edges.parent.data(), sizeof(Index) * ne, h);
  h = fnv(mesh.edges.kids.data(), sizeof(Index) * 2 * ne, h);
  std::printf("%016llx\n", (unsigned long long)h);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int depth = std::atoi(argv[2]), amr = std::atoi(argv[3]), passes = std::atoi(argv[4]);
  try {
    const std::string seed(argv[1]);
    if (seed == "icos") return run<IcosTriSphereSeed>(depth, amr, passes);
    if (seed == "quad_rect") return run<QuadRectSeed>(depth, amr, passes);
    if (seed == "tri_hex") return run<TriHexSeed>(depth, amr, passes);
    return run<CubedSphereSeed>(depth, amr, passes);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 4;
  }
}
