// lpm/lpm_coords.hpp, fields -- Coords<Geo> (src/lpm_coords.hpp:37-220) and ScalarField / VectorField
// (src/lpm_field.hpp:30-144): a `view`, its `hview` host mirror (the same host storage here), a name and units.
#ifndef LPM_SHIM_COORDS_HPP
#define LPM_SHIM_COORDS_HPP

#include <map>
#include <utility>

#include "lpm_geometry.hpp"

namespace Lpm {

template <typename Geo>
struct Coords {
  typedef typename Geo::crd_view_type view_type;
  view_type view;
  view_type hview;
  Coords() = default;
  explicit Coords(const Index nmax) : view("crds", nmax), hview(view), nmax_(nmax), n_(0) {}
  Index nh() const { return n_; }
  Index n_max() const { return nmax_; }
  void set_nh(const Index n) { n_ = n; }
  void update_device() const {}
  void update_host() const {}

 private:
  Index nmax_ = 0, n_ = 0;
};

enum FieldLocation { ParticleField, VertexField, EdgeField, FaceField };
inline std::string field_loc_string(const FieldLocation& fl) {
  switch (fl) {
    case ParticleField: return "particle_field";
    case VertexField: return "vertex_field";
    case EdgeField: return "edge_field";
    default: return "face_field";
  }
}

template <FieldLocation FL>
struct ScalarField {
  typedef scalar_view_type view_type;
  typedef std::map<std::string, std::string> metadata_type;
  static constexpr FieldLocation field_loc = FL;
  static constexpr int ndim = 1;
  scalar_view_type view;
  scalar_view_type hview;
  std::string name;
  std::string units;
  metadata_type metadata;
  ScalarField() = default;
  ScalarField(const std::string& mname, const Index nmax, const std::string& u = "null_unit")
      : view(mname, nmax), hview(view), name(mname), units(u) {
    metadata.emplace("name", mname);
    metadata.emplace("location", field_loc_string(FL));
    metadata.emplace("units", u);
  }
  Real operator()(const Index i) const { return view(i); }
  void update_device() const {}
  void update_host() const {}
  /// (min, max) over the first n values (src/lpm_field.hpp: range)
  std::pair<Real, Real> range(const Index n) const {
    Real lo = view(0), hi = view(0);
    for (Index i = 1; i < n; ++i) {
      lo = std::min(lo, view(i));
      hi = std::max(hi, view(i));
    }
    return std::make_pair(lo, hi);
  }
  bool has_nan(const Index n) const {
    for (Index i = 0; i < n; ++i)
      if (std::isnan(view(i))) return true;
    return false;
  }
};

template <typename Geo, FieldLocation FL>
struct VectorField {
  typedef typename Geo::vec_view_type view_type;
  static constexpr FieldLocation field_loc = FL;
  static constexpr int ndim = Geo::ndim;
  view_type view;
  view_type hview;
  std::string name;
  std::string units;
  VectorField() = default;
  VectorField(const std::string& mname, const Index nmax, const std::string& u = "null_unit")
      : view(mname, nmax), hview(view), name(mname), units(u) {}
  void update_device() const {}
  void update_host() const {}
  /// (min, max) of the vector magnitude over the first n entries
  std::pair<Real, Real> range(const Index n) const {
    Real lo = 0, hi = 0;
    for (Index i = 0; i < n; ++i) {
      const Real m = Geo::mag(view.row(i));
      if (i == 0) lo = hi = m;
      lo = std::min(lo, m);
      hi = std::max(hi, m);
    }
    return std::make_pair(lo, hi);
  }
};

}  // namespace Lpm
#endif
