// lpm/lpm_vtk_io.hpp -- VtkPolymeshInterface<SeedType> (src/vtk/lpm_vtk_io.hpp:33-97, _impl.hpp:13-263) without
// libvtk: the same data set written as a VTK XML PolyData (.vtp) file by hand.
//   points      every vertex, phys_crds (z = 0 in the plane)                          make_points     _impl.hpp:49-62
//   polys       the leaf faces only (has_kids == false), in face order, nfaceverts each  make_cells  _impl.hpp:92-104
//   cell data   "area" (leaves), then "lag_crds"; add_*_cell_data compacts to the leaves  _impl.hpp:121-134,198-260
//   point data  "lag_crds"; add_*_point_data takes every vertex                       _impl.hpp:146-196
//   array names name.empty() ? view.label() : name                                    _impl.hpp:159
// vtkXMLPolyDataWriter's default encoding (appended, zlib, base64) needs libvtk/zlib; this writer emits the equally
// valid format="ascii" arrays with 17 significant digits, so a reader recovers every double bit-exactly.  The file is
// the same VTK data set, not the same bytes.
// Deviation, flagged: the reference's height-field constructor never inserts a point on the sphere
// (_impl.hpp:78-86 computes (1 + h) x and drops it -- the file would have no points); here the evident intent is written.
#ifndef LPM_SHIM_VTK_IO_HPP
#define LPM_SHIM_VTK_IO_HPP

#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "lpm_polymesh2d.hpp"

namespace Lpm {

template <typename SeedType>
class VtkPolymeshInterface {
 public:
  using geo = typename SeedType::geo;

  explicit VtkPolymeshInterface(const PolyMesh2d<SeedType>& pm) : mesh_(pm) {
    make_points(nullptr);
    init_common();
  }
  VtkPolymeshInterface(const PolyMesh2d<SeedType>& pm, const scalar_view_type height_field) : mesh_(pm) {
    make_points(&height_field);
    init_common();
  }

  void update_positions() { make_points(nullptr); }

  template <typename VT = scalar_view_type>
  void add_scalar_point_data(const VT s, const std::string& name = "") {
    Array a{name.empty() ? s.label() : name, 1, {}};
    a.v.reserve(mesh_.n_vertices_host());
    for (Index i = 0; i < mesh_.n_vertices_host(); ++i) a.v.push_back(s(i));
    point_data_.push_back(std::move(a));
  }
  template <typename VT = typename geo::vec_view_type>
  void add_vector_point_data(const VT v, const std::string& name = "") {
    const int nc = (int)v.extent(1);
    Array a{name.empty() ? v.label() : name, nc, {}};
    for (Index i = 0; i < mesh_.n_vertices_host(); ++i)
      for (int k = 0; k < nc; ++k) a.v.push_back(v(i, k));
    point_data_.push_back(std::move(a));
  }
  template <typename VT = scalar_view_type>
  void add_scalar_cell_data(const VT s, const std::string& name = "") {
    Array a{name.empty() ? s.label() : name, 1, {}};
    for (Index i = 0; i < mesh_.n_faces_host(); ++i)
      if (!mesh_.faces.mask(i)) a.v.push_back(s(i));
    cell_data_.push_back(std::move(a));
  }
  template <typename VT = typename geo::vec_view_type>
  void add_vector_cell_data(const VT v, const std::string& name = "") {
    const int nc = (int)v.extent(1);
    Array a{name.empty() ? v.label() : name, nc, {}};
    for (Index i = 0; i < mesh_.n_faces_host(); ++i)
      if (!mesh_.faces.mask(i))
        for (int k = 0; k < nc; ++k) a.v.push_back(v(i, k));
    cell_data_.push_back(std::move(a));
  }
  void add_tracers(const std::vector<scalar_view_type>& point_tracers, const std::vector<scalar_view_type>& cell_tracers) {
    LPM_REQUIRE_MSG(point_tracers.size() == cell_tracers.size(), "add_tracers: vertex and face tracer counts differ");
    for (size_t k = 0; k < point_tracers.size(); ++k) {
      add_scalar_point_data(point_tracers[k]);
      add_scalar_cell_data(cell_tracers[k]);
    }
  }

  void write(const std::string& ofilename) const {
    std::ofstream os(ofilename);
    LPM_REQUIRE_MSG(os.good(), "VtkPolymeshInterface::write: cannot open " + ofilename);
    const Index npts = (Index)(points_.size() / 3);
    const Index ncells = n_leaves();
    os << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <PolyData>\n"
       << "    <Piece NumberOfPoints=\"" << npts << "\" NumberOfVerts=\"0\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\""
       << ncells << "\">\n";
    os << "      <PointData>\n";
    for (const auto& a : point_data_) write_array(os, a);
    os << "      </PointData>\n      <CellData>\n";
    for (const auto& a : cell_data_) write_array(os, a);
    os << "      </CellData>\n      <Points>\n";
    write_array(os, Array{"Points", 3, points_});
    os << "      </Points>\n      <Polys>\n        <DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n";
    constexpr int nfv = SeedType::faceKind::nverts;
    for (Index i = 0; i < mesh_.n_faces_host(); ++i) {
      if (mesh_.faces.mask(i)) continue;
      os << "         ";
      for (int j = 0; j < nfv; ++j) os << " " << mesh_.faces.verts(i, j);
      os << "\n";
    }
    os << "        </DataArray>\n        <DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n         ";
    for (Index c = 1; c <= ncells; ++c) os << " " << (long)c * nfv << (c % 16 == 0 ? "\n         " : "");
    os << "\n        </DataArray>\n      </Polys>\n    </Piece>\n  </PolyData>\n</VTKFile>\n";
  }

 protected:
  struct Array {
    std::string name;
    int ncomp;
    std::vector<Real> v;
  };
  const PolyMesh2d<SeedType>& mesh_;
  std::vector<Real> points_;
  std::vector<Array> point_data_, cell_data_;

  Index n_leaves() const {
    Index n = 0;
    for (Index i = 0; i < mesh_.n_faces_host(); ++i) n += mesh_.faces.mask(i) ? 0 : 1;
    return n;
  }
  void init_common() {
    add_scalar_cell_data(mesh_.faces.area, "area");
    add_vector_point_data(mesh_.vertices.lag_crds.view, "lag_crds");
    add_vector_cell_data(mesh_.faces.lag_crds.view, "lag_crds");
  }
  void make_points(const scalar_view_type* height) {
    points_.clear();
    const auto x = mesh_.vertices.phys_crds.view;
    for (Index i = 0; i < mesh_.n_vertices_host(); ++i) {
      if (geo::ndim == 3) {
        const Real s = height ? 1 + (*height)(i) : 1;
        for (int k = 0; k < 3; ++k) points_.push_back(height ? s * x(i, k) : x(i, k));
      } else {
        points_.push_back(x(i, 0));
        points_.push_back(x(i, 1));
        points_.push_back(height ? (*height)(i) : 0);
      }
    }
  }
  static void write_array(std::ostream& os, const Array& a) {
    os << "        <DataArray type=\"Float64\" Name=\"" << a.name << "\" NumberOfComponents=\"" << a.ncomp << "\" format=\"ascii\">\n";
    char buf[32];
    for (size_t i = 0; i < a.v.size(); ++i) {
      if (i % a.ncomp == 0) os << (i ? "\n          " : "          ");
      std::snprintf(buf, sizeof(buf), "%.17g", a.v[i]);
      os << buf << " ";
    }
    os << "\n        </DataArray>\n";
  }
};

}  // namespace Lpm
#endif
