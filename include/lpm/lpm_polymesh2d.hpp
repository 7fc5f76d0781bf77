// lpm/lpm_polymesh2d.hpp -- MeshSeed<Seed>, PolyMeshParameters<Seed> and PolyMesh2d<Seed>, backed by the C ABI's host mesh
// generator (lpmx_mesh_*); the two spherical seeds are defined here, the planar ones in lpm_plane.hpp.
//   MeshSeed / IcosTriSphereSeed / CubedSphereSeed     src/mesh/lpm_mesh_seed.hpp:111-147,150-230
//   MeshSeed::set_max_allocations                      src/mesh/lpm_mesh_seed.cpp:266-279
//   PolyMeshParameters                                 src/mesh/lpm_polymesh2d.hpp:32-72
//   PolyMesh2d::tree_init / n_*_host / appx_mesh_size  src/mesh/lpm_polymesh2d.hpp:152-159,861; _impl.hpp:25-42
//   Vertices / Edges / Faces public views              src/mesh/lpm_vertices.hpp, lpm_edges.hpp, lpm_faces.hpp
//   PolyMesh2d::divide_flagged_faces                   src/mesh/lpm_polymesh2d_impl.hpp:124-173 (via lpmx_mesh_divide_flagged_faces)
// As in the reference every mesh view is allocated once at its maximum extent (nmaxverts / nmaxedges / nmaxfaces, sized
// for init_depth + amr_buffer) and n_*_host() counts the entries in use, so shallow copies held by flag functors and
// solvers stay valid across an adaptive refinement.  The mesh queries (ccw_adjacent_faces, locate_face_containing_pt, ...) forward to
// the host generator (lpmx_mesh_*), which is where the mesh lives.
#ifndef LPM_SHIM_POLYMESH2D_HPP
#define LPM_SHIM_POLYMESH2D_HPP

#include <limits>
#include <memory>

#include "lpm_coords.hpp"
#include "lpm_logger.hpp"

namespace Lpm {

struct TriFace {
  static constexpr Int nverts = 3;
};
struct QuadFace {
  static constexpr Int nverts = 4;
};

struct CubedSphereSeed {
  static constexpr Int nverts = 8, nfaces = 6, nedges = 12, nfaceverts = 4, vertex_degree = 4;
  static constexpr int lpmx_id = LPMX_SEED_CUBED_SPHERE;
  typedef SphereGeometry geo;
  typedef QuadFace faceKind;
  static std::string filename() { return "cubedSphereSeed.dat"; }
  static std::string id_string() { return "cubed_sphere"; }
};

struct IcosTriSphereSeed {
  static constexpr Int nverts = 12, nfaces = 20, nedges = 30, nfaceverts = 3, vertex_degree = 6;
  static constexpr int lpmx_id = LPMX_SEED_ICOS_TRI_SPHERE;
  typedef SphereGeometry geo;
  typedef TriFace faceKind;
  static std::string filename() { return "icosTriSphereSeed.dat"; }
  static std::string id_string() { return "icostri_sphere"; }
};

template <typename SeedType>
struct MeshSeed {
  Real radius;
  explicit MeshSeed(const Real r = 1) : radius(r) {}
  static std::string id_string() { return SeedType::id_string(); }
  /// memory needed for a uniform tree of depth lev (src/mesh/lpm_mesh_seed.cpp:266-279)
  void set_max_allocations(Index& nboundary, Index& nedges, Index& nfaces, const Int lev) const {
    LPM_REQUIRE(lpmx_mesh_max_allocations(SeedType::lpmx_id, lev, &nboundary, &nedges, &nfaces) == LPMX_OK);
  }
};

template <typename SeedType>
struct PolyMeshParameters {
  Index nmaxverts, nmaxedges, nmaxfaces;
  Int init_depth, amr_buffer, amr_limit;
  MeshSeed<SeedType> seed;
  PolyMeshParameters() : nmaxverts(0), nmaxedges(0), nmaxfaces(0), init_depth(0), amr_buffer(0), amr_limit(0), seed() {}
  PolyMeshParameters(const Int depth, const Real r = 1, const Int amr_buff = 0, const Int amr_lim = 0)
      : init_depth(depth), amr_buffer(amr_buff), amr_limit(amr_lim), seed(r) {
    seed.set_max_allocations(nmaxverts, nmaxedges, nmaxfaces, depth + amr_buff);
  }
};

template <typename Geo>
struct Vertices {
  Coords<Geo> phys_crds, lag_crds;
  index_view_type crd_inds;
  Index nh() const { return n_; }
  Index n_ = 0;
};

struct Edges {
  index_view_type origs, dests, lefts, rights, parent;
  View2<Index, 2> kids;
  Index nh() const { return n_; }
  Index n_leaves_host() const { return n_leaves_; }
  Index n_ = 0, n_leaves_ = 0;
};

template <typename FaceKind, typename Geo>
struct Faces {
  Coords<Geo> phys_crds, lag_crds;
  scalar_view_type area;
  mask_view_type mask;              ///< non-zero = divided panel (not a leaf)
  View2<Index, FaceKind::nverts> verts, edges;
  View2<Index, 4> kids;
  index_view_type crd_inds, parent, level, leaf_idx;
  Index nh() const { return n_; }
  Index n_leaves_host() const { return n_leaves_; }
  /// sqrt(mean leaf area) (src/mesh/lpm_faces_impl.hpp:213-226)
  Real appx_mesh_size() const {
    Real s = 0;
    for (Index i = 0; i < n_; ++i)
      if (!mask(i)) s += area(i);
    return std::sqrt(s / n_leaves_);
  }
  /// sqrt(smallest leaf area) (src/mesh/lpm_faces_impl.hpp:228-240; Kokkos::Min identity = largest double)
  Real appx_min_mesh_size() const {
    Real result = std::numeric_limits<Real>::max();
    for (Index i = 0; i < n_; ++i)
      if (!mask(i)) result = (area(i) < result ? area(i) : result);
    return std::sqrt(result);
  }
  Real surface_area_host() const {
    Real s = 0;
    for (Index i = 0; i < n_; ++i) s += area(i);
    return s;
  }
  Index n_ = 0, n_leaves_ = 0;
};

template <typename SeedType>
class PolyMesh2d {
 public:
  typedef SeedType seed_type;
  typedef typename SeedType::geo Geo;
  typedef typename SeedType::faceKind FaceType;

  Vertices<Geo> vertices;
  Edges edges;
  Faces<FaceType, Geo> faces;
  Real radius = 1;
  Int base_tree_depth = 0;

  PolyMeshParameters<SeedType> params;  ///< init_depth / amr_buffer / amr_limit / nmax* (src/mesh/lpm_polymesh2d.hpp:137)

  /// allocates; tree_init() fills (reference: PolyMesh2d(nmaxverts, nmaxedges, nmaxfaces))
  PolyMesh2d(const Index nmaxverts, const Index nmaxedges, const Index nmaxfaces)
      : nmaxverts_(nmaxverts), nmaxedges_(nmaxedges), nmaxfaces_(nmaxfaces) {
    params.nmaxverts = nmaxverts, params.nmaxedges = nmaxedges, params.nmaxfaces = nmaxfaces;
  }

  /// allocates and builds the uniform tree (reference: PolyMesh2d(const PolyMeshParameters&), lpm_polymesh2d.hpp:152-159)
  explicit PolyMesh2d(const PolyMeshParameters<SeedType>& params_in)
      : params(params_in), nmaxverts_(params_in.nmaxverts), nmaxedges_(params_in.nmaxedges), nmaxfaces_(params_in.nmaxfaces) {
    tree_init(params.init_depth, params.seed);
  }
  virtual ~PolyMesh2d() = default;

  /// PolyMesh2d::tree_init (src/mesh/lpm_polymesh2d_impl.hpp:25-42) via lpmx_mesh_create
  void tree_init(const Int initDepth, const MeshSeed<SeedType>& seed) {
    lpmx_mesh_t m = nullptr;
    LPM_REQUIRE_MSG(lpmx_mesh_create(SeedType::lpmx_id, initDepth, seed.radius, &m) == LPMX_OK, "lpmx_mesh_create");
    handle_ = std::shared_ptr<lpmx_mesh_s>(m, [](lpmx_mesh_s* p) { lpmx_mesh_destroy(p); });
    int nv, ne, nf;
    lpmx_mesh_sizes(m, &nv, &ne, &nf, nullptr, nullptr, nullptr);
    LPM_REQUIRE_MSG(nv <= nmaxverts_ && ne <= nmaxedges_ && nf <= nmaxfaces_, "mesh exceeds the allocated sizes");
    radius = seed.radius;
    base_tree_depth = initDepth;
    params.init_depth = initDepth;
    const Index mv = nmaxverts_, me = nmaxedges_, mf = nmaxfaces_;
    vertices.phys_crds = Coords<Geo>(mv), vertices.lag_crds = Coords<Geo>(mv);
    faces.phys_crds = Coords<Geo>(mf), faces.lag_crds = Coords<Geo>(mf);
    vertices.crd_inds = index_view_type("vert_crd_inds", mv);
    edges.origs = index_view_type("origs", me), edges.dests = index_view_type("dests", me);
    edges.lefts = index_view_type("lefts", me), edges.rights = index_view_type("rights", me);
    edges.parent = index_view_type("edge_parent", me), edges.kids = View2<Index, 2>("edge_kids", me);
    faces.area = scalar_view_type("area", mf), faces.mask = mask_view_type("mask", mf);
    faces.verts = View2<Index, FaceType::nverts>("face_verts", mf), faces.edges = View2<Index, FaceType::nverts>("face_edges", mf);
    faces.kids = View2<Index, 4>("face_kids", mf);
    faces.crd_inds = index_view_type("face_crd_inds", mf), faces.parent = index_view_type("face_parent", mf);
    faces.level = index_view_type("face_level", mf), faces.leaf_idx = index_view_type("leaf_idx", mf);
    refetch();
  }

  /// PolyMesh2d::divide_flagged_faces (src/mesh/lpm_polymesh2d_impl.hpp:124-173).  The particles may have moved since the
  /// mesh was built: the current coordinates are handed to the generator first (the reference divides from the
  /// coordinates as they are), every view is refilled in place afterwards.
  template <typename LoggerType>
  void divide_flagged_faces(const mask_view_type& flags, LoggerType& logger) {
    LPM_REQUIRE_MSG(handle_, "divide_flagged_faces before tree_init");
    LPM_REQUIRE((Index)flags.extent(0) >= n_faces_host());
    Index flag_count = 0;
    for (Index i = 0; i < n_faces_host(); ++i) flag_count += (flags(i) ? 1 : 0);
    logger.debug("dividing {} flagged faces...", flag_count);
    const long nd = Geo::ndim;
    push(LPMX_MESH_VERT_XYZ, vertices.phys_crds.view.data(), nd * n_vertices_host());
    push(LPMX_MESH_VERT_LAG_XYZ, vertices.lag_crds.view.data(), nd * n_vertices_host());
    push(LPMX_MESH_FACE_XYZ, faces.phys_crds.view.data(), nd * n_faces_host());
    push(LPMX_MESH_FACE_LAG_XYZ, faces.lag_crds.view.data(), nd * n_faces_host());
    int refine_count = 0, outcome = 0;
    const int rc = lpmx_mesh_divide_flagged_faces(handle_.get(), flags.data(), (int)flags.extent(0), nmaxfaces_,
                                                  params.init_depth + params.amr_limit, &refine_count, &outcome);
    LPM_REQUIRE_MSG(rc == LPMX_OK, std::string("lpmx_mesh_divide_flagged_faces: ") + lpmx_error_name(rc));
    if (outcome == LPMX_AMR_NO_SPACE) {
      logger.warn("divide_flagged_faces: not enough memory (flag count = {}, nfaces = {}, nmaxfaces = {})", flag_count,
                  n_faces_host(), nmaxfaces_);
      return;
    }
    refetch();
    if (outcome == LPMX_AMR_LIMIT_REACHED)
      logger.warn("divide_flagged_faces: local refinement limit reached; divided {} of {} flagged faces.", refine_count,
                  flag_count);
    else
      logger.info("divide_flagged_faces: {} faces divided.", refine_count);
  }

  // ---- mesh queries (src/mesh/lpm_polymesh2d.hpp:262-552; host code here as the mesh is) ----
  /// PolyMesh2d::get_leaf_edges_from_parent (:277-308)
  template <typename EdgeList = Index*>
  void get_leaf_edges_from_parent(EdgeList& edge_list, Int& n_leaves, const Index parent_edge_idx) const {
    query(&lpmx_mesh_leaf_edges_from_parent, parent_edge_idx, edge_list, n_leaves, "get_leaf_edges_from_parent");
  }
  /// PolyMesh2d::ccw_edges_around_face (:316-340)
  template <typename EdgeList = Index*>
  void ccw_edges_around_face(EdgeList& face_leaf_edges, Int& n_leaf_edges, const Index face_idx) const {
    query(&lpmx_mesh_ccw_edges_around_face, face_idx, face_leaf_edges, n_leaf_edges, "ccw_edges_around_face");
  }
  /// PolyMesh2d::ccw_adjacent_faces (:348-366); LPM_NULL_IDX across a free boundary
  template <typename FaceList = Index*>
  void ccw_adjacent_faces(FaceList& adj_faces, Int& n_adj, const Index face_idx) const {
    query(&lpmx_mesh_ccw_adjacent_faces, face_idx, adj_faces, n_adj, "ccw_adjacent_faces");
  }
  /// NeighborsFlag::operator() over faces [start, end) (src/mesh/lpm_refinement_flags.hpp:30-52)
  void neighbors_flag(const mask_view_type& flags, const Index start, const Index end) const {
    LPM_REQUIRE((Index)flags.extent(0) >= end);
    LPM_REQUIRE(lpmx_mesh_neighbors_flag(handle_.get(), flags.data(), (int)start, (int)end, nullptr) == LPMX_OK);
  }
  /// The generator answers point-location queries from ITS coordinates: hand it the views' current values first.
  void push_coordinates() const {
    const long nd = Geo::ndim;
    lpmx_mesh_t m = handle_.get();
    LPM_REQUIRE(lpmx_mesh_update_array(m, LPMX_MESH_VERT_XYZ, vertices.phys_crds.view.data(), nd * n_vertices_host()) == LPMX_OK);
    LPM_REQUIRE(lpmx_mesh_update_array(m, LPMX_MESH_FACE_XYZ, faces.phys_crds.view.data(), nd * n_faces_host()) == LPMX_OK);
  }
  /// PolyMesh2d::locate_face_containing_pt (:541-552), locate_pt_walk_search (:377-413), locate_pt_tree_search (:446-473),
  /// nearest_root_face (:421-434).  Call push_coordinates() after the particles have moved.
  template <typename Point>
  Index locate_face_containing_pt(const Point& query_pt) const {
    return locate(LPMX_LOCATE_CONTAINING, query_pt, 0);
  }
  template <typename Point>
  Index locate_pt_walk_search(const Point& query_pt, const Index face_start_idx) const {
    return locate(LPMX_LOCATE_WALK, query_pt, face_start_idx);
  }
  template <typename Point>
  Index locate_pt_tree_search(const Point& query_pt, const Index root_face) const {
    return locate(LPMX_LOCATE_TREE, query_pt, root_face);
  }
  template <typename Point>
  Index nearest_root_face(const Point& query_pt) const {
    return locate(LPMX_LOCATE_NEAREST_ROOT, query_pt, 0);
  }

  Index n_vertices_host() const { return vertices.nh(); }
  Index n_edges_host() const { return edges.nh(); }
  Index n_faces_host() const { return faces.nh(); }
  Real appx_mesh_size() const { return faces.appx_mesh_size(); }
  Real appx_min_mesh_size() const { return faces.appx_min_mesh_size(); }
  Real surface_area_host() const { return faces.surface_area_host(); }
  virtual void update_device() const {}
  virtual void update_host() const {}

  virtual std::string info_string(const std::string& label = "", const int tab_level = 0, const bool = false) const {
    std::ostringstream ss;
    const std::string tabs(tab_level, '\t');
    ss << tabs << "PolyMesh2d<" << SeedType::id_string() << "> " << label << ": depth " << base_tree_depth << ", "
       << n_vertices_host() << " vertices, " << n_edges_host() << " edges (" << edges.n_leaves_host() << " leaves), "
       << n_faces_host() << " faces (" << faces.n_leaves_host() << " leaves); surface area " << surface_area_host()
       << ", appx mesh size " << appx_mesh_size() << "\n";
    return ss.str();
  }

 protected:
  Index nmaxverts_, nmaxedges_, nmaxfaces_;

 private:
  std::shared_ptr<lpmx_mesh_s> handle_;  // the host generator's tree (kept for divide_flagged_faces)

  template <typename List>
  void query(int (*fn)(lpmx_mesh_t, int, int*, int, int*), const Index idx, List& list, Int& n, const char* what) const {
    int tmp[8 * 6 * 4];  // nfaceverts * LPM_MAX_AMR_LIMIT entries in the reference's callers; generous
    int nn = 0;
    LPM_REQUIRE_MSG(fn(handle_.get(), (int)idx, tmp, (int)(sizeof(tmp) / sizeof(tmp[0])), &nn) == LPMX_OK, what);
    LPM_REQUIRE_MSG(nn <= (int)(sizeof(tmp) / sizeof(tmp[0])), what);
    for (int i = 0; i < nn; ++i) list[i] = tmp[i];
    n = nn;
  }
  template <typename Point>
  Index locate(const int mode, const Point& pt, const Index start) const {
    Real q[3] = {0, 0, 0};
    for (int k = 0; k < Geo::ndim; ++k) q[k] = pt[k];
    const int st = (int)start;
    int out = -1;
    LPM_REQUIRE(lpmx_mesh_locate(handle_.get(), mode, q, 1, &st, &out) == LPMX_OK);
    return out;
  }

  void push(const int id, const Real* src, const long count) {
    LPM_REQUIRE(lpmx_mesh_update_array(handle_.get(), id, src, count) == LPMX_OK);
  }

  // counts + every array, into the nmax-sized views
  void refetch() {
    lpmx_mesh_t m = handle_.get();
    int nv, ne, nf, nfl, nel, nfv;
    lpmx_mesh_sizes(m, &nv, &ne, &nf, &nfl, &nel, &nfv);
    LPM_REQUIRE_MSG(nv <= nmaxverts_ && ne <= nmaxedges_ && nf <= nmaxfaces_, "mesh exceeds the allocated sizes");
    vertices.n_ = nv, edges.n_ = ne, edges.n_leaves_ = nel, faces.n_ = nf, faces.n_leaves_ = nfl;
    vertices.phys_crds.set_nh(nv), vertices.lag_crds.set_nh(nv), faces.phys_crds.set_nh(nf), faces.lag_crds.set_nh(nf);
    fetch(m, LPMX_MESH_VERT_XYZ, vertices.phys_crds.view.data());
    fetch(m, LPMX_MESH_VERT_LAG_XYZ, vertices.lag_crds.view.data());
    fetch(m, LPMX_MESH_VERT_CRD_INDS, vertices.crd_inds.data());
    fetch(m, LPMX_MESH_EDGE_ORIGS, edges.origs.data());
    fetch(m, LPMX_MESH_EDGE_DESTS, edges.dests.data());
    fetch(m, LPMX_MESH_EDGE_LEFTS, edges.lefts.data());
    fetch(m, LPMX_MESH_EDGE_RIGHTS, edges.rights.data());
    fetch(m, LPMX_MESH_EDGE_PARENTS, edges.parent.data());
    fetch(m, LPMX_MESH_EDGE_KIDS, edges.kids.data());
    fetch(m, LPMX_MESH_FACE_XYZ, faces.phys_crds.view.data());
    fetch(m, LPMX_MESH_FACE_LAG_XYZ, faces.lag_crds.view.data());
    fetch(m, LPMX_MESH_FACE_AREA, faces.area.data());
    fetch(m, LPMX_MESH_FACE_MASK, faces.mask.data());
    fetch(m, LPMX_MESH_FACE_VERTS, faces.verts.data());
    fetch(m, LPMX_MESH_FACE_EDGES, faces.edges.data());
    fetch(m, LPMX_MESH_FACE_CRD_INDS, faces.crd_inds.data());
    fetch(m, LPMX_MESH_FACE_PARENT, faces.parent.data());
    fetch(m, LPMX_MESH_FACE_KIDS, faces.kids.data());
    fetch(m, LPMX_MESH_FACE_LEVEL, faces.level.data());
    fetch(m, LPMX_MESH_FACE_LEAF_IDX, faces.leaf_idx.data());
  }

  template <typename T>
  static void fetch(lpmx_mesh_t m, const int id, T* dst) {
    const void* src = nullptr;
    long n = 0;
    int kind = 0;
    LPM_REQUIRE(lpmx_mesh_array(m, id, &src, &n, &kind) == LPMX_OK);
    const T* s = static_cast<const T*>(src);
    for (long i = 0; i < n; ++i) dst[i] = s[i];
  }
};

}  // namespace Lpm
#endif
