// lpmx_mesh.cpp -- host-side quad-tree particle/panel mesh generator for the two spherical seeds and the two planar seeds
// with free boundaries (QuadRectSeed, TriHexSeed; PlaneGeometry, Real*[2] coordinates).
//
// What it reproduces (reference files under /root/reference/src):
//   MeshSeed<Seed>                          mesh/lpm_mesh_seed.cpp:10-18 (radius), :266-279 (allocations)
//   PolyMesh2d<Seed>::seed_init/tree_init   mesh/lpm_polymesh2d_impl.hpp:18-42
//   Vertices::init_from_seed                mesh/lpm_vertices_impl.hpp:32-40
//   Edges::init_from_seed / divide          mesh/lpm_edges.cpp:42-96
//   Faces::init_from_seed / insert_host     mesh/lpm_faces_impl.hpp:126-203
//   FaceDivider<Sphere,TriFace>::divide     mesh/lpm_faces_impl.hpp:284-431
//   FaceDivider<Sphere,QuadFace>::divide    mesh/lpm_faces_impl.hpp:433-574
//   SphereGeometry midpoint/barycenter/tri_area/polygon_area   lpm_geometry.hpp:516-642
//   PlaneGeometry  midpoint/barycenter/tri_area/polygon_area   lpm_geometry.hpp:69-160
//   Faces::scan_leaves                      mesh/lpm_faces_impl.hpp:100-124
//
// The design is a flat structure-of-arrays builder with the reference's insertion order (that order
// is what fixes every integer array); nothing here is Kokkos-shaped.  Floating-point expressions
// are evaluated exactly as written, and this file is compiled with -ffp-contract=off so the
// coordinates do not depend on the compiler's FMA contraction choices.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/lpmx.h"

namespace {

#include "seed_tables.inc"

constexpr int kNull = -1;  // constants::NULL_IND
constexpr double kZeroTol = 2.220446049250313e-16;  // FloatingPoint<Real>::zero_tol (lpm_floating_point.hpp:22)

struct Vec3 {
  double v[3];
};

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(double* c, const double* a, const double* b) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double mag3(const double* a) { return std::sqrt(dot3(a, a)); }
inline void scale3(double s, double* a) {
  a[0] *= s;
  a[1] *= s;
  a[2] *= s;
}
// SphereGeometry::normalize (lpm_geometry.hpp:449-452): scale by 1/|v|, not divide
inline void normalize3(double* a) { scale3(1.0 / mag3(a), a); }
// SphereGeometry::distance (:459-465)
inline double sph_dist(const double* a, const double* b) {
  double cp[3];
  cross3(cp, a, b);
  return std::atan2(mag3(cp), dot3(a, b));
}
// SphereGeometry::tri_area (:593-608)
inline double tri_area(const double* a, const double* b, const double* c) {
  const double s1 = sph_dist(a, b);
  const double s2 = sph_dist(b, c);
  const double s3 = sph_dist(c, a);
  const double half_perim = 0.5 * (s1 + s2 + s3);
  double zz = std::tan(0.5 * half_perim) * std::tan(0.5 * (half_perim - s1)) *
              std::tan(0.5 * (half_perim - s2)) * std::tan(0.5 * (half_perim - s3));
  if (std::fabs(zz) < kZeroTol) zz = 0;
  return 4 * std::atan(std::sqrt(zz));
}
// SphereGeometry::polygon_area (:632-642)
inline double polygon_area(const double* ctr, const double (*verts)[3], int n) {
  double ar = 0;
  for (int i = 0; i < n; ++i) ar += tri_area(ctr, verts[i], verts[(i + 1) % n]);
  return ar;
}
// SphereGeometry::barycenter (:524-535)
inline void barycenter(double* out, const double (*verts)[3], int n) {
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < n; ++i) {
    out[0] += verts[i][0];
    out[1] += verts[i][1];
    out[2] += verts[i][2];
  }
  scale3(1.0 / n, out);
  normalize3(out);
}
// SphereGeometry::midpoint (:578-584)
inline void midpoint(double* out, const double* a, const double* b) {
  out[0] = 0.5 * (a[0] + b[0]);
  out[1] = 0.5 * (a[1] + b[1]);
  out[2] = 0.5 * (a[2] + b[2]);
  normalize3(out);
}

// ---- PlaneGeometry (lpm_geometry.hpp:69-160); the third component of the local arrays is carried as 0 and never read ----
inline double plane_tri_area(const double* a, const double* b, const double* c) {
  const double bma0 = b[0] - a[0], bma1 = b[1] - a[1];
  const double cma0 = c[0] - a[0], cma1 = c[1] - a[1];
  const double ar = bma0 * cma1 - bma1 * cma0;
  return 0.5 * std::fabs(ar);
}
inline double plane_polygon_area(const double* ctr, const double (*verts)[3], int n) {
  double ar = 0.0;
  for (int i = 0; i < n; ++i) ar += plane_tri_area(ctr, verts[i], verts[(i + 1) % n]);
  return ar;
}
inline void plane_barycenter(double* out, const double (*verts)[3], int n) {
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < n; ++i) {
    out[0] += verts[i][0];
    out[1] += verts[i][1];
  }
  const double s = 1.0 / n;
  out[0] *= s;
  out[1] *= s;
}
inline void plane_midpoint(double* out, const double* a, const double* b) {
  out[0] = 0.5 * (a[0] + b[0]);
  out[1] = 0.5 * (a[1] + b[1]);
  out[2] = 0.0;
}

long ipow4(int lev) {
  long r = 1;
  for (int i = 0; i < lev; ++i) r *= 4;
  return r;
}

}  // namespace

struct lpmx_mesh_s {
  int seed = 0;
  int nfv = 3;  // vertices per face
  int nd = 3;   // Geo::ndim: 3 on the sphere, 2 in the plane (row length of every coordinate array)
  int depth = 0;
  // vertices
  std::vector<double> vx, vlag;  // [nv][nd]
  std::vector<int> v_crd;
  // edges
  std::vector<int> eo, ed, el, er, ep, ek;  // ek: [ne][2]
  int edge_leaves = 0;
  // faces
  std::vector<double> fx, flag, farea;  // fx/flag: [nf][nd]
  std::vector<unsigned char> fmask;
  std::vector<int> fverts, fedges;  // [nf][nfv]
  std::vector<int> f_crd, fparent, fkids, flevel, fleaf;  // fkids: [nf][4]
  int face_leaves = 0;

  int nv() const { return (int)v_crd.size(); }
  int ne() const { return (int)eo.size(); }
  int nf() const { return (int)f_crd.size(); }

  int insert_vertex(const double* p, const double* l) {
    const int idx = nv();
    vx.insert(vx.end(), p, p + nd);
    vlag.insert(vlag.end(), l, l + nd);
    v_crd.push_back(idx);  // crd index == insertion index (lpm_vertices.hpp:158-163)
    return idx;
  }
  int insert_edge(int o, int d, int l, int r, int prt) {
    const int idx = ne();
    eo.push_back(o);
    ed.push_back(d);
    el.push_back(l);
    er.push_back(r);
    ep.push_back(prt);
    ek.push_back(kNull);
    ek.push_back(kNull);
    ++edge_leaves;
    return idx;
  }
  bool edge_has_kids(int e) const { return ek[2 * e] > 0; }  // lpm_edges.hpp:290-293
  bool face_has_kids(int f) const { return fkids[4 * f] > 0; }  // lpm_faces.hpp:270-273
  int insert_face(const double* p, const double* l, const int* verts, const int* edges, int prt, double ar) {
    const int idx = nf();
    fx.insert(fx.end(), p, p + nd);
    flag.insert(flag.end(), l, l + nd);
    fverts.insert(fverts.end(), verts, verts + nfv);
    fedges.insert(fedges.end(), edges, edges + nfv);
    for (int i = 0; i < 4; ++i) fkids.push_back(kNull);
    f_crd.push_back(idx);
    fparent.push_back(prt);
    farea.push_back(ar);
    // Faces::insert_host reads level(parent)+1 even for parent == NULL_IND (an out-of-bounds read
    // in the reference, zero in practice): root level is DEFINED as 1 here (SURVEY.md Mesh-ii).
    flevel.push_back(prt == kNull ? 1 : flevel[prt] + 1);
    fmask.push_back(0);
    ++face_leaves;
    return idx;
  }

  // Geo::midpoint / barycenter / polygon_area on 3-wide local rows
  void geo_midpoint(double* out, const double* a, const double* b) const {
    if (nd == 3) midpoint(out, a, b);
    else plane_midpoint(out, a, b);
  }
  void geo_barycenter(double* out, const double (*verts)[3], int n) const {
    if (nd == 3) barycenter(out, verts, n);
    else plane_barycenter(out, verts, n);
  }
  double geo_polygon_area(const double* ctr, const double (*verts)[3], int n) const {
    return nd == 3 ? polygon_area(ctr, verts, n) : plane_polygon_area(ctr, verts, n);
  }
  // row `idx` of a coordinate array as a 3-wide local (z = 0 in the plane)
  void load_row(double* out, const std::vector<double>& a, int idx) const {
    out[2] = 0.0;
    for (int k = 0; k < nd; ++k) out[k] = a[(size_t)nd * idx + k];
  }

  // Edges::divide (lpm_edges.cpp:58-96); returns index of first child
  int divide_edge(int e) {
    const int vins = nv();
    const int eins = ne();
    double a[3], b[3], la[3], lb[3];
    load_row(a, vx, v_crd[eo[e]]);
    load_row(b, vx, v_crd[ed[e]]);
    // lag_endpts(1,:) is filled from phys_crds in the reference (:81); identical at build time.
    load_row(la, vlag, v_crd[eo[e]]);
    load_row(lb, vx, v_crd[ed[e]]);
    double mid[3], lmid[3];
    geo_midpoint(mid, a, b);
    geo_midpoint(lmid, la, lb);
    insert_vertex(mid, lmid);
    const int o = eo[e], d = ed[e], l = el[e], r = er[e];
    insert_edge(o, vins, l, r, e);
    insert_edge(vins, d, l, r, e);
    ek[2 * e] = eins;
    ek[2 * e + 1] = eins + 1;
    --edge_leaves;
    return eins;
  }

  void divide_tri(int f);
  void divide_quad(int f);
  // Faces::scan_leaves: exclusive scan of !has_kids
  void scan_leaves() {
    fleaf.resize(nf());
    int psum = 0;
    for (int i = 0; i < nf(); ++i) {
      fleaf[i] = psum;
      psum += face_has_kids(i) ? 0 : 1;
    }
  }
  void finish_parent(int f, int first_kid) {
    for (int i = 0; i < 4; ++i) fkids[4 * f + i] = first_kid + i;
    farea[f] = 0.0;
    fmask[f] = 1;
    --face_leaves;
  }
};

// FaceDivider<SphereGeometry, TriFace>::divide (lpm_faces_impl.hpp:284-431)
void lpmx_mesh_s::divide_tri(int f) {
  int nfe[4][3], nfvx[4][3];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) nfe[i][j] = nfvx[i][j] = kNull;
  const int pv[3] = {fverts[3 * f], fverts[3 * f + 1], fverts[3 * f + 2]};
  const int pe[3] = {fedges[3 * f], fedges[3 * f + 1], fedges[3 * f + 2]};
  const int kid0 = nf();
  for (int i = 0; i < 3; ++i) nfvx[i][i] = pv[i];
  for (int i = 0; i < 3; ++i) {
    const int pedge = pe[i];
    int k0, k1;
    if (edge_has_kids(pedge)) {
      k0 = ek[2 * pedge];
      k1 = ek[2 * pedge + 1];
    } else {
      k0 = divide_edge(pedge);
      k1 = k0 + 1;
    }
    const int a = i, b = (i + 1) % 3;
    if (f == el[pedge]) {  // positive orientation (lpm_faces.hpp:359-362)
      nfe[a][i] = k0;
      el[k0] = kid0 + a;
      nfe[b][i] = k1;
      el[k1] = kid0 + b;
    } else {
      nfe[a][i] = k1;
      er[k1] = kid0 + a;
      nfe[b][i] = k0;
      er[k0] = kid0 + b;
    }
    const int m = ed[k0];  // the new midpoint vertex
    if (i == 0) {
      nfvx[0][1] = m;
      nfvx[1][0] = m;
      nfvx[3][2] = m;
    } else if (i == 1) {
      nfvx[1][2] = m;
      nfvx[2][1] = m;
      nfvx[3][0] = m;
    } else {
      nfvx[2][0] = m;
      nfvx[0][2] = m;
      nfvx[3][1] = m;
    }
  }
  // three interior edges, all with the centre kid (3) on the left
  const int e0 = ne();
  for (int i = 0; i < 3; ++i) nfe[3][i] = e0 + i;
  nfe[0][1] = e0 + 1;
  nfe[1][2] = e0 + 2;
  nfe[2][0] = e0;
  insert_edge(nfvx[2][1], nfvx[2][0], kid0 + 3, kid0 + 2, kNull);
  insert_edge(nfvx[0][2], nfvx[0][1], kid0 + 3, kid0 + 0, kNull);
  insert_edge(nfvx[1][0], nfvx[1][2], kid0 + 3, kid0 + 1, kNull);
  // kid centres and areas are all computed before any kid is appended
  double ctr[4][3], lctr[4][3], area[4];
  for (int i = 0; i < 4; ++i) {
    double vc[3][3], vl[3][3];
    for (int j = 0; j < 3; ++j) {
      load_row(vc[j], vx, v_crd[nfvx[i][j]]);
      load_row(vl[j], vlag, v_crd[nfvx[i][j]]);
    }
    geo_barycenter(ctr[i], vc, 3);
    geo_barycenter(lctr[i], vl, 3);
    area[i] = geo_polygon_area(ctr[i], vc, 3);
  }
  for (int i = 0; i < 4; ++i) insert_face(ctr[i], lctr[i], nfvx[i], nfe[i], f, area[i]);
  finish_parent(f, kid0);
}

// FaceDivider<SphereGeometry, QuadFace>::divide (lpm_faces_impl.hpp:433-574)
void lpmx_mesh_s::divide_quad(int f) {
  int nfe[4][4], nfvx[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) nfe[i][j] = nfvx[i][j] = kNull;
  int pv[4], pe[4];
  for (int i = 0; i < 4; ++i) {
    pv[i] = fverts[4 * f + i];
    pe[i] = fedges[4 * f + i];
  }
  const int kid0 = nf();
  for (int i = 0; i < 4; ++i) nfvx[i][i] = pv[i];
  for (int i = 0; i < 4; ++i) {
    const int pedge = pe[i];
    int k0, k1;
    if (edge_has_kids(pedge)) {
      k0 = ek[2 * pedge];
      k1 = ek[2 * pedge + 1];
    } else {
      k0 = divide_edge(pedge);
      k1 = k0 + 1;
    }
    const int a = i, b = (i + 1) % 4;
    if (f == el[pedge]) {
      nfe[a][i] = k0;
      el[k0] = kid0 + a;
      nfe[b][i] = k1;
      el[k1] = kid0 + b;
    } else {
      nfe[a][i] = k1;
      er[k1] = kid0 + a;
      nfe[b][i] = k0;
      er[k0] = kid0 + b;
    }
    const int m = ed[k0];
    nfvx[a][b] = m;
    nfvx[b][a] = m;
  }
  // the parent's centre becomes a vertex, appended after the edge midpoints (:506-517)
  const int pc = f_crd[f];
  double pcx[3], pcl[3];
  load_row(pcx, fx, pc);
  load_row(pcl, flag, pc);
  const int cv = insert_vertex(pcx, pcl);
  for (int i = 0; i < 4; ++i) nfvx[i][(i + 2) % 4] = cv;
  // four interior edges (:520-536)
  const int e0 = ne();
  insert_edge(nfvx[0][1], nfvx[0][2], kid0 + 0, kid0 + 1, kNull);
  nfe[0][1] = e0;
  nfe[1][3] = e0;
  insert_edge(nfvx[2][0], nfvx[2][3], kid0 + 3, kid0 + 2, kNull);
  nfe[2][3] = e0 + 1;
  nfe[3][1] = e0 + 1;
  insert_edge(nfvx[2][1], nfvx[2][0], kid0 + 1, kid0 + 2, kNull);
  nfe[1][2] = e0 + 2;
  nfe[2][0] = e0 + 2;
  insert_edge(nfvx[3][1], nfvx[3][0], kid0 + 0, kid0 + 3, kNull);
  nfe[0][2] = e0 + 3;
  nfe[3][0] = e0 + 3;
  double ctr[4][3], lctr[4][3], area[4];
  for (int i = 0; i < 4; ++i) {
    double vc[4][3], vl[4][3];
    for (int j = 0; j < 4; ++j) {
      // the quad divider indexes coordinates by vertex id directly (:548-552)
      load_row(vc[j], vx, nfvx[i][j]);
      load_row(vl[j], vlag, nfvx[i][j]);
    }
    geo_barycenter(ctr[i], vc, 4);
    geo_barycenter(lctr[i], vl, 4);
    area[i] = geo_polygon_area(ctr[i], vc, 4);
  }
  for (int i = 0; i < 4; ++i) insert_face(ctr[i], lctr[i], nfvx[i], nfe[i], f, area[i]);
  finish_parent(f, kid0);
}

namespace {

struct SeedDesc {
  int nverts, nedges, nfaces, nfv, nd;
  const double (*crds)[3];
  const int (*edges)[4];
  const int* fverts;
  const int* fedges;
};

bool seed_desc(int seed, SeedDesc* d) {
  if (seed == LPMX_SEED_ICOS_TRI_SPHERE) {
    *d = {12, 30, 20, 3, 3, kIcosTri_crds, kIcosTri_edges, &kIcosTri_face_verts[0][0], &kIcosTri_face_edges[0][0]};
    return true;
  }
  if (seed == LPMX_SEED_CUBED_SPHERE) {
    *d = {8, 12, 6, 4, 3, kCubedSphere_crds, kCubedSphere_edges, &kCubedSphere_face_verts[0][0],
          &kCubedSphere_face_edges[0][0]};
    return true;
  }
  if (seed == LPMX_SEED_QUAD_RECT) {
    *d = {9, 12, 4, 4, 2, kQuadRect_crds, kQuadRect_edges, &kQuadRect_face_verts[0][0], &kQuadRect_face_edges[0][0]};
    return true;
  }
  if (seed == LPMX_SEED_TRI_HEX) {
    *d = {7, 12, 6, 3, 2, kTriHex_crds, kTriHex_edges, &kTriHex_face_verts[0][0], &kTriHex_face_edges[0][0]};
    return true;
  }
  return false;
}

// Seed::n_vertices_at_tree_level / n_faces_at_tree_level / n_edges_at_tree_level (lpm_mesh_seed.cpp:306-375)
long nverts_at(int seed, int lev) {
  if (seed == LPMX_SEED_QUAD_RECT) {  // (3 + 2 + 4 + ... + 2^lev)^2
    long r = 3;
    for (int i = 1; i <= lev; ++i) r += 1L << i;
    return r * r;
  }
  if (seed == LPMX_SEED_TRI_HEX) {  // 2 * sum_{i = 2^lev + 1}^{2^(lev+1)} i + 2^(lev+1) + 1
    long r = 0;
    for (long i = (1L << lev) + 1; i <= (1L << (lev + 1)); ++i) r += i;
    return 2 * r + (1L << (lev + 1)) + 1;
  }
  return 2 + (seed == LPMX_SEED_ICOS_TRI_SPHERE ? 10 : 6) * ipow4(lev);
}
long nfaces_at(int seed, int lev) {
  const int n0 = seed == LPMX_SEED_ICOS_TRI_SPHERE ? 20 : seed == LPMX_SEED_QUAD_RECT ? 4 : 6;
  return n0 * ipow4(lev);
}
// Euler: V - E + F = 2 on the sphere, 1 for the planar seeds (free boundary)
long nedges_at(int seed, int lev) {
  const bool planar = seed == LPMX_SEED_QUAD_RECT || seed == LPMX_SEED_TRI_HEX;
  return nverts_at(seed, lev) + nfaces_at(seed, lev) - (planar ? 1 : 2);
}

// ------------------------------------------------------------------------------------------------
// Mesh queries (src/mesh/lpm_polymesh2d.hpp:262-552), restated as coded.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxAmrLimit = 6;  // LPM_MAX_AMR_LIMIT (LpmConfig.h.in:19): bounds the reference's fixed-size index lists

// PolyMesh2d::get_leaf_edges_from_parent (:277-308).  The reference makes room for the second kid with an ASCENDING copy
// (edge_list[j + 1] = edge_list[j], j = i + 1 ...), which smears edge_list[i + 1] over every later entry when more than one
// entry follows the divided edge (an edge refined three levels deeper on one side): replicated, the list is the contract.
// The reference's list has 2 * LPM_MAX_AMR_LIMIT slots and no bounds check; this one grows instead of overflowing.
void leaf_edges_from_parent(const lpmx_mesh_s& m, int parent, std::vector<int>& list) {
  list.assign(1, parent);
  int n_leaves = 1;
  bool keep_going = m.edge_has_kids(parent);
  while (keep_going) {
    int n_new = 0;
    keep_going = false;
    for (int i = 0; i < n_leaves; ++i) {
      if ((size_t)(n_leaves + 2) > list.size()) list.resize(n_leaves + 2, kNull);
      if (m.edge_has_kids(list[i])) {
        const int kid0 = m.ek[2 * list[i]], kid1 = m.ek[2 * list[i] + 1];
        for (int j = i + 1; j < n_leaves; ++j) list[j + 1] = list[j];
        list[i] = kid0;
        list[i + 1] = kid1;
        if (m.edge_has_kids(kid0) || m.edge_has_kids(kid1)) keep_going = true;
        ++n_new;
      }
    }
    n_leaves += n_new;
  }
  list.resize(n_leaves);
}

// PolyMesh2d::edge_is_positive (:262-265)
inline bool edge_is_positive(const lpmx_mesh_s& m, int e, int f) { return f == m.el[e]; }

// PolyMesh2d::ccw_edges_around_face (:316-340)
void ccw_edges_around_face(const lpmx_mesh_s& m, int f, std::vector<int>& out) {
  out.clear();
  std::vector<int> leaves;
  for (int i = 0; i < m.nfv; ++i) {
    const int e = m.fedges[(size_t)m.nfv * f + i];
    leaf_edges_from_parent(m, e, leaves);
    if (edge_is_positive(m, e, f))
      out.insert(out.end(), leaves.begin(), leaves.end());
    else
      out.insert(out.end(), leaves.rbegin(), leaves.rend());
  }
}

// PolyMesh2d::ccw_adjacent_faces (:348-366); NULL_IND (-1) across a free boundary
void ccw_adjacent_faces(const lpmx_mesh_s& m, int f, std::vector<int>& out) {
  std::vector<int> le;
  ccw_edges_around_face(m, f, le);
  out.clear();
  for (int e : le) out.push_back(edge_is_positive(m, e, f) ? m.er[e] : m.el[e]);
}

// Geo::distance: great-circle angle (lpm_geometry.hpp:470-475) / Euclidean (:62-67)
inline double geo_distance(const lpmx_mesh_s& m, const double* a, const double* b) {
  if (m.nd == 3) return sph_dist(a, b);
  const double d0 = b[0] - a[0], d1 = b[1] - a[1];
  return std::sqrt(d0 * d0 + d1 * d1);
}

// Geo::barycenter of a face's vertices at their current coordinates
void face_vertex_barycenter(const lpmx_mesh_s& m, int f, double* out) {
  double vs[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < m.nfv; ++k) {
    const int v = m.fverts[(size_t)m.nfv * f + k];
    for (int c = 0; c < m.nd; ++c) vs[k][c] = m.vx[(size_t)m.nd * v + c];
  }
  m.geo_barycenter(out, vs, m.nfv);
}

// PolyMesh2d::nearest_root_face (:421-434): strict '<', first root face wins ties
int nearest_root_face(const lpmx_mesh_s& m, const double* q, int n_roots) {
  int result = 0;
  double dist = geo_distance(m, q, &m.fx[0]);
  for (int i = 1; i < n_roots; ++i) {
    const double t = geo_distance(m, &m.fx[(size_t)m.nd * i], q);
    if (t < dist) dist = t, result = i;
  }
  return result;
}

// PolyMesh2d::locate_pt_tree_search (:446-473): descend to the kid whose centre is closest
int locate_pt_tree_search(const lpmx_mesh_s& m, const double* q, int root) {
  int cur = root;
  while (m.face_has_kids(cur)) {
    int next = m.fkids[4 * (size_t)cur];
    double dist = geo_distance(m, &m.fx[(size_t)m.nd * next], q);
    for (int k = 1; k < 4; ++k) {
      const int kid = m.fkids[4 * (size_t)cur + k];
      const double t = geo_distance(m, q, &m.fx[(size_t)m.nd * kid]);
      if (t < dist) next = kid, dist = t;
    }
    cur = next;
  }
  return cur;
}

// PolyMesh2d::pt_is_outside_mesh (:484-535), planar meshes only: closer to the mirror image of the face centroid across one
// of the face's boundary edges than to the centroid itself
bool pt_is_outside_mesh(const lpmx_mesh_s& m, const double* q, int f) {
  if (m.nd != 2) return false;
  std::vector<int> le;
  ccw_edges_around_face(m, f, le);
  bool any = false;
  for (int e : le) any = any || m.el[e] == kNull || m.er[e] == kNull;
  if (!any) return false;
  double ctr[3];
  face_vertex_barycenter(m, f, ctr);
  const double intr = geo_distance(m, q, ctr);
  bool result = false;
  for (int e : le) {
    if (!(m.el[e] == kNull || m.er[e] == kNull)) continue;
    const double* o = &m.vx[2 * (size_t)m.eo[e]];
    const double* d = &m.vx[2 * (size_t)m.ed[e]];
    double qv[2] = {d[0] - o[0], d[1] - o[1]};
    const double* v0 = o;
    if (!edge_is_positive(m, e, f)) {
      qv[0] = -qv[0], qv[1] = -qv[1];
      v0 = d;
    }
    const double len = std::sqrt(qv[0] * qv[0] + qv[1] * qv[1]);  // PlaneGeometry::normalize (:45-50): scale by 1 / |v|
    qv[0] *= 1.0 / len, qv[1] *= 1.0 / len;
    const double pv[2] = {ctr[0] - v0[0], ctr[1] - v0[1]};
    const double dotp = pv[0] * qv[0] + pv[1] * qv[1];
    const double refl[3] = {ctr[0] - 2 * (pv[0] - dotp * qv[0]), ctr[1] - 2 * (pv[1] - dotp * qv[1]), 0.0};
    if (geo_distance(m, q, refl) < intr) result = true;
  }
  return result;
}

// PolyMesh2d::locate_pt_walk_search (:377-413): the starting distance is to the face's stored centre, the neighbours are
// judged by the barycentre of their vertices
int locate_pt_walk_search(const lpmx_mesh_s& m, const double* q, int start) {
  int result = start, cur = start;
  double dist = geo_distance(m, q, &m.fx[(size_t)m.nd * cur]);
  std::vector<int> adj;
  bool keep_going = true;
  while (keep_going) {
    result = cur;
    ccw_adjacent_faces(m, cur, adj);
    for (int a : adj) {
      if (a == kNull) continue;
      double ctr[3];
      face_vertex_barycenter(m, a, ctr);
      const double t = geo_distance(m, ctr, q);
      if (t < dist) dist = t, cur = a;
    }
    keep_going = cur != result;
  }
  return result;
}

}  // namespace

extern "C" {

int lpmx_mesh_leaf_edges_from_parent(lpmx_mesh_t m, int parent_edge, int* list, int cap, int* n) {
  if (!m || !n || parent_edge < 0 || parent_edge >= m->ne() || cap < 0 || (cap > 0 && !list)) return LPMX_ERR_INVALID;
  std::vector<int> v;
  leaf_edges_from_parent(*m, parent_edge, v);
  *n = (int)v.size();
  for (int i = 0; i < *n && i < cap; ++i) list[i] = v[i];
  return LPMX_OK;
}

int lpmx_mesh_ccw_edges_around_face(lpmx_mesh_t m, int face, int* list, int cap, int* n) {
  if (!m || !n || face < 0 || face >= m->nf() || cap < 0 || (cap > 0 && !list)) return LPMX_ERR_INVALID;
  std::vector<int> v;
  ccw_edges_around_face(*m, face, v);
  *n = (int)v.size();
  for (int i = 0; i < *n && i < cap; ++i) list[i] = v[i];
  return LPMX_OK;
}

int lpmx_mesh_ccw_adjacent_faces(lpmx_mesh_t m, int face, int* list, int cap, int* n) {
  if (!m || !n || face < 0 || face >= m->nf() || cap < 0 || (cap > 0 && !list)) return LPMX_ERR_INVALID;
  std::vector<int> v;
  ccw_adjacent_faces(*m, face, v);
  *n = (int)v.size();
  for (int i = 0; i < *n && i < cap; ++i) list[i] = v[i];
  return LPMX_OK;
}

// NeighborsFlag::operator() (src/mesh/lpm_refinement_flags.hpp:30-52) over faces [start, end)
int lpmx_mesh_neighbors_flag(lpmx_mesh_t m, unsigned char* flags, int start, int end, int* n_flagged) {
  if (!m || !flags || start < 0 || end < start || end > m->nf()) return LPMX_ERR_INVALID;
  std::vector<int> adj;
  int count = 0;
  for (int i = start; i < end; ++i) {
    ccw_adjacent_faces(*m, i, adj);
    const int lev = m->flevel[i];
    bool refine = false;
    for (int a : adj) {
      // across a free boundary the reference reads faces.level(-1) (out of bounds); no neighbour, no level here
      if (a != kNull && m->flevel[a] > lev + 1) {
        refine = true;
        break;
      }
    }
    if (refine && !flags[i]) ++count;
    flags[i] = (flags[i] || refine) ? 1 : 0;
  }
  if (n_flagged) *n_flagged = count;
  return LPMX_OK;
}

int lpmx_mesh_locate(lpmx_mesh_t m, int mode, const double* pts, int n_pts, const int* start, int* out) {
  if (!m || n_pts < 0 || (n_pts > 0 && (!pts || !out)) || m->nf() == 0) return LPMX_ERR_INVALID;
  if ((mode == LPMX_LOCATE_WALK || mode == LPMX_LOCATE_TREE) && n_pts > 0 && !start) return LPMX_ERR_INVALID;
  SeedDesc d;
  if (!seed_desc(m->seed, &d)) return LPMX_ERR_INVALID;
  for (int i = 0; i < n_pts; ++i) {
    const double* q = pts + (size_t)m->nd * i;
    switch (mode) {
      case LPMX_LOCATE_CONTAINING: {  // PolyMesh2d::locate_face_containing_pt (:541-552)
        const int leaf = locate_pt_tree_search(*m, q, nearest_root_face(*m, q, d.nfaces));
        out[i] = pt_is_outside_mesh(*m, q, leaf) ? kNull : locate_pt_walk_search(*m, q, leaf);
        break;
      }
      case LPMX_LOCATE_WALK:
        if (start[i] < 0 || start[i] >= m->nf() || m->face_has_kids(start[i])) return LPMX_ERR_INVALID;  // leaf-only (:382)
        out[i] = locate_pt_walk_search(*m, q, start[i]);
        break;
      case LPMX_LOCATE_TREE:
        if (start[i] < 0 || start[i] >= m->nf()) return LPMX_ERR_INVALID;
        out[i] = locate_pt_tree_search(*m, q, start[i]);
        break;
      case LPMX_LOCATE_NEAREST_ROOT:
        out[i] = nearest_root_face(*m, q, d.nfaces);
        break;
      default:
        return LPMX_ERR_INVALID;
    }
  }
  return LPMX_OK;
}

int lpmx_mesh_max_allocations(int seed, int depth, int* n_verts, int* n_edges, int* n_faces) {
  SeedDesc d;
  if (!seed_desc(seed, &d) || depth < 0 || depth > 12 || !n_verts || !n_edges || !n_faces) return LPMX_ERR_INVALID;
  long nv = nverts_at(seed, depth), ne = 0, nf = 0;
  for (int i = 0; i <= depth; ++i) {
    nf += nfaces_at(seed, i);
    ne += nedges_at(seed, i);
  }
  if (nf > 2000000000L || ne > 2000000000L) return LPMX_ERR_UNSUPPORTED;  // Index is int in the reference
  *n_verts = (int)nv;
  *n_edges = (int)ne;
  *n_faces = (int)nf;
  return LPMX_OK;
}

int lpmx_mesh_create(int seed, int depth, double radius, lpmx_mesh_t* out) {
  SeedDesc d;
  if (!out || !seed_desc(seed, &d) || depth < 0 || !(radius > 0)) return LPMX_ERR_INVALID;
  int nvmax, nemax, nfmax;
  const int rc = lpmx_mesh_max_allocations(seed, depth, &nvmax, &nemax, &nfmax);
  if (rc != LPMX_OK) return rc;
  lpmx_mesh_s* m = new (std::nothrow) lpmx_mesh_s;
  if (!m) return LPMX_ERR_NOMEM;
  try {
    m->seed = seed;
    m->nfv = d.nfv;
    m->nd = d.nd;
    m->depth = depth;
    m->vx.reserve((long)d.nd * nvmax);
    m->vlag.reserve((long)d.nd * nvmax);
    m->v_crd.reserve(nvmax);
    for (auto* v : {&m->eo, &m->ed, &m->el, &m->er, &m->ep}) v->reserve(nemax);
    m->ek.reserve(2L * nemax);
    m->fx.reserve((long)d.nd * nfmax);
    m->flag.reserve((long)d.nd * nfmax);
    m->farea.reserve(nfmax);
    m->fmask.reserve(nfmax);
    m->fverts.reserve((long)d.nfv * nfmax);
    m->fedges.reserve((long)d.nfv * nfmax);
    m->fkids.reserve(4L * nfmax);
    for (auto* v : {&m->f_crd, &m->fparent, &m->flevel}) v->reserve(nfmax);

    // MeshSeed(maxr): all seed coordinates are multiplied by the radius (lpm_mesh_seed.cpp:10-18)
    std::vector<Vec3> sc(d.nverts + d.nfaces);
    for (int i = 0; i < d.nverts + d.nfaces; ++i)
      for (int k = 0; k < 3; ++k) sc[i].v[k] = (radius == 1.0) ? d.crds[i][k] : d.crds[i][k] * radius;
    // seed_init: vertices, edges, faces in that order
    for (int i = 0; i < d.nverts; ++i) m->insert_vertex(sc[i].v, sc[i].v);
    for (int i = 0; i < d.nedges; ++i) m->insert_edge(d.edges[i][0], d.edges[i][1], d.edges[i][2], d.edges[i][3], kNull);
    for (int i = 0; i < d.nfaces; ++i) {
      double vc[4][3];
      for (int j = 0; j < d.nfv; ++j)
        for (int k = 0; k < 3; ++k) vc[j][k] = sc[d.fverts[i * d.nfv + j]].v[k];
      const double ar = m->geo_polygon_area(sc[d.nverts + i].v, vc, d.nfv);  // MeshSeed::face_area (:281-294)
      m->insert_face(sc[d.nverts + i].v, sc[d.nverts + i].v, &d.fverts[i * d.nfv], &d.fedges[i * d.nfv], kNull, ar);
    }
    // tree_init (lpm_polymesh2d_impl.hpp:25-42), including startInd = stopInd - 1
    int start = 0;
    for (int lev = 0; lev < depth; ++lev) {
      const int stop = m->nf();
      for (int j = start; j < stop; ++j) {
        if (!m->face_has_kids(j)) {
          if (d.nfv == 3)
            m->divide_tri(j);
          else
            m->divide_quad(j);
        }
      }
      start = stop - 1;
    }
    m->scan_leaves();
  } catch (const std::bad_alloc&) {
    delete m;
    return LPMX_ERR_NOMEM;
  }
  *out = m;
  return LPMX_OK;
}

int lpmx_mesh_destroy(lpmx_mesh_t mesh) {
  delete mesh;
  return LPMX_OK;
}

int lpmx_mesh_sizes(lpmx_mesh_t m, int* n_verts, int* n_edges, int* n_faces, int* n_face_leaves,
                    int* n_edge_leaves, int* n_face_verts) {
  if (!m) return LPMX_ERR_INVALID;
  if (n_verts) *n_verts = m->nv();
  if (n_edges) *n_edges = m->ne();
  if (n_faces) *n_faces = m->nf();
  if (n_face_leaves) *n_face_leaves = m->face_leaves;
  if (n_edge_leaves) *n_edge_leaves = m->edge_leaves;
  if (n_face_verts) *n_face_verts = m->nfv;
  return LPMX_OK;
}

int lpmx_mesh_update_array(lpmx_mesh_t m, int id, const double* data, long count) {
  if (!m || !data) return LPMX_ERR_INVALID;
  std::vector<double>* dst = nullptr;
  switch (id) {
    case LPMX_MESH_VERT_XYZ: dst = &m->vx; break;
    case LPMX_MESH_VERT_LAG_XYZ: dst = &m->vlag; break;
    case LPMX_MESH_FACE_XYZ: dst = &m->fx; break;
    case LPMX_MESH_FACE_LAG_XYZ: dst = &m->flag; break;
    default: return LPMX_ERR_INVALID;
  }
  if (count != (long)dst->size()) return LPMX_ERR_INVALID;
  std::copy(data, data + count, dst->begin());
  return LPMX_OK;
}

// PolyMesh2d::divide_flagged_faces (lpm_polymesh2d_impl.hpp:124-173)
int lpmx_mesh_divide_flagged_faces(lpmx_mesh_t m, const unsigned char* flags, int n_flags, int max_faces, int max_level,
                                   int* n_divided, int* outcome) {
  if (!m || (!flags && n_flags > 0) || n_flags < 0) return LPMX_ERR_INVALID;
  const int n_in = m->nf();
  if (n_flags < n_in) return LPMX_ERR_INVALID;
  int flag_count = 0;
  for (int i = 0; i < n_in; ++i) {
    if (flags[i]) {
      if (m->face_has_kids(i)) return LPMX_ERR_INVALID;
      ++flag_count;
    }
  }
  if (n_divided) *n_divided = 0;
  const int space_left = max_faces - n_in;
  if (flag_count > space_left / 4) {  // "not enough memory": warn and return, nothing divided (:138-144)
    if (outcome) *outcome = LPMX_AMR_NO_SPACE;
    return LPMX_OK;
  }
  int refine_count = 0;
  bool limit_reached = false;
  try {
    for (int i = 0; i < n_in; ++i) {
      if (!flags[i]) continue;
      if (m->flevel[i] <= max_level) {
        if (m->nfv == 3)
          m->divide_tri(i);
        else
          m->divide_quad(i);
        ++refine_count;
      } else {
        limit_reached = true;
      }
    }
    m->scan_leaves();
  } catch (const std::bad_alloc&) {
    return LPMX_ERR_NOMEM;
  }
  if (n_divided) *n_divided = refine_count;
  if (outcome) *outcome = limit_reached ? LPMX_AMR_LIMIT_REACHED : LPMX_AMR_DIVIDED_ALL;
  return LPMX_OK;
}

int lpmx_mesh_array(lpmx_mesh_t m, int id, const void** data, long* count, int* is_real) {
  if (!m || !data || !count) return LPMX_ERR_INVALID;
  int kind = 0;
  const void* p = nullptr;
  long n = 0;
#define RET_D(vec) p = (vec).data(), n = (long)(vec).size(), kind = 1
#define RET_I(vec) p = (vec).data(), n = (long)(vec).size(), kind = 0
  switch (id) {
    case LPMX_MESH_VERT_XYZ: RET_D(m->vx); break;
    case LPMX_MESH_VERT_LAG_XYZ: RET_D(m->vlag); break;
    case LPMX_MESH_VERT_CRD_INDS: RET_I(m->v_crd); break;
    case LPMX_MESH_EDGE_ORIGS: RET_I(m->eo); break;
    case LPMX_MESH_EDGE_DESTS: RET_I(m->ed); break;
    case LPMX_MESH_EDGE_LEFTS: RET_I(m->el); break;
    case LPMX_MESH_EDGE_RIGHTS: RET_I(m->er); break;
    case LPMX_MESH_EDGE_PARENTS: RET_I(m->ep); break;
    case LPMX_MESH_EDGE_KIDS: RET_I(m->ek); break;
    case LPMX_MESH_FACE_XYZ: RET_D(m->fx); break;
    case LPMX_MESH_FACE_LAG_XYZ: RET_D(m->flag); break;
    case LPMX_MESH_FACE_AREA: RET_D(m->farea); break;
    case LPMX_MESH_FACE_MASK: p = m->fmask.data(), n = (long)m->fmask.size(), kind = 2; break;
    case LPMX_MESH_FACE_VERTS: RET_I(m->fverts); break;
    case LPMX_MESH_FACE_EDGES: RET_I(m->fedges); break;
    case LPMX_MESH_FACE_CRD_INDS: RET_I(m->f_crd); break;
    case LPMX_MESH_FACE_PARENT: RET_I(m->fparent); break;
    case LPMX_MESH_FACE_KIDS: RET_I(m->fkids); break;
    case LPMX_MESH_FACE_LEVEL: RET_I(m->flevel); break;
    case LPMX_MESH_FACE_LEAF_IDX: RET_I(m->fleaf); break;
    default: return LPMX_ERR_INVALID;
  }
#undef RET_D
#undef RET_I
  *data = p;
  *count = n;
  if (is_real) *is_real = kind;
  return LPMX_OK;
}

}  // extern "C"
