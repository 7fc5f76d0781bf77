// lpm/lpm_views.hpp -- host-resident stand-ins for the Kokkos Views of the reference's public members.
//
// Semantics mirrored from Kokkos (SURVEY.md 8(b) "value semantics"): zero-initialised on allocation, reference
// counted shallow copies, operator() element access that is callable on const objects, extent(), data().
//   scalar_view_type = View<Real*>, mask_view_type = View<bool*>, index_view_type = View<Index*>
//                                                    src/lpm_kokkos_defs.hpp:31-39
//   crd_view_type / vec_view_type = View<Real*[3]>   src/lpm_geometry.hpp:261-263  (LayoutRight, as on the host)
#ifndef LPM_SHIM_VIEWS_HPP
#define LPM_SHIM_VIEWS_HPP

#include <memory>
#include <string>
#include <vector>

#include "lpm_config.hpp"

namespace Lpm {

template <typename T>
class View1 {
 public:
  typedef View1<T> HostMirror;
  static constexpr int rank = 1;
  View1() : d_(std::make_shared<std::vector<T>>()) {}
  View1(const std::string& label, const size_t n) : label_(label), d_(std::make_shared<std::vector<T>>(n, T(0))) {}
  T& operator()(const size_t i) const { return (*d_)[i]; }
  T& operator[](const size_t i) const { return (*d_)[i]; }
  size_t extent(const int) const { return d_->size(); }
  size_t size() const { return d_->size(); }
  T* data() const { return d_->data(); }
  const std::string& label() const { return label_; }

 private:
  std::string label_;
  std::shared_ptr<std::vector<T>> d_;
};

/// A row of a rank-2 view: what Kokkos::subview(v, i, Kokkos::ALL) returns.
template <typename T>
struct RowRef {
  T* p;
  T& operator()(const int j) const { return p[j]; }
  T& operator[](const int j) const { return p[j]; }
};

template <typename T, int N>
class View2 {
 public:
  typedef View2<T, N> HostMirror;
  static constexpr int rank = 2;
  View2() : d_(std::make_shared<std::vector<T>>()) {}
  View2(const std::string& label, const size_t n) : label_(label), d_(std::make_shared<std::vector<T>>(n * N, T(0))) {}
  T& operator()(const size_t i, const int j) const { return (*d_)[i * N + j]; }
  RowRef<T> row(const size_t i) const { return RowRef<T>{d_->data() + i * N}; }
  size_t extent(const int dim) const { return dim == 0 ? d_->size() / N : N; }
  T* data() const { return d_->data(); }
  const std::string& label() const { return label_; }

 private:
  std::string label_;
  std::shared_ptr<std::vector<T>> d_;
};

typedef View1<Real> scalar_view_type;
typedef View1<Index> index_view_type;
typedef View1<unsigned char> mask_view_type;  // Kokkos View<bool*>: one byte per entry, non-zero = divided panel
typedef View2<Real, 3> vec3_view_type;

namespace ko {
struct ALL_t {};
static constexpr ALL_t ALL{};
template <typename T, int N>
inline RowRef<T> subview(const View2<T, N>& v, const size_t i, ALL_t) { return v.row(i); }
template <typename V>
inline V create_mirror_view(const V& v) { return v; }
template <typename V>
inline void deep_copy(const V& dst, const V& src) {
  if (dst.data() != src.data())
    for (size_t i = 0; i < dst.extent(0) * (V::rank == 2 ? dst.extent(1) : 1); ++i) dst.data()[i] = src.data()[i];
}
}  // namespace ko

}  // namespace Lpm
#endif
