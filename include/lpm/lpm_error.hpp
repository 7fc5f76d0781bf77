// lpm/lpm_error.hpp -- ErrNorms: weighted l1 / l2 / linf error norms (src/lpm_error.hpp:81-131,
// src/lpm_error_impl.hpp:59-108).  l1 = sum|e|w / sum|exact|w, l2 = sqrt(sum e^2 w / sum exact^2 w),
// linf = max|e| / max|exact|; for rank-2 views |.| is the Euclidean magnitude of a row.  Divided panels carry
// area 0, so they drop out of l1/l2 but enter the unmasked linf, as in the reference.
#ifndef LPM_SHIM_ERROR_HPP
#define LPM_SHIM_ERROR_HPP

#include <iomanip>

#include "lpm_geometry.hpp"

namespace Lpm {

struct ENormScalar {
  Real l1num = 0, l1denom = 0, l2num = 0, l2denom = 0, linfnum = 0, linfdenom = 0;
};

namespace impl {
inline Real row_mag(const scalar_view_type& v, const Index i) { return std::abs(v(i)); }
inline Real row_mag(const vec3_view_type& v, const Index i) { return SphereGeometry::mag(v.row(i)); }
inline void set_err(const scalar_view_type& e, const scalar_view_type& a, const scalar_view_type& x, const Index i, const bool m) {
  e(i) = m ? 0 : a(i) - x(i);
}
inline void set_err(const vec3_view_type& e, const vec3_view_type& a, const vec3_view_type& x, const Index i, const bool m) {
  for (int j = 0; j < 3; ++j) e(i, j) = m ? 0 : a(i, j) - x(i, j);
}
}  // namespace impl

struct ErrNorms {
  Real l1, l2, linf;
  ErrNorms(const Real l_1, const Real l_2, const Real l_i) : l1(l_1), l2(l_2), linf(l_i) {}
  explicit ErrNorms(const ENormScalar& err)
      : l1(err.l1num / err.l1denom), l2(std::sqrt(err.l2num / err.l2denom)), linf(err.linfnum / err.linfdenom) {}

  /// computes err = appx - exact, then reduces
  template <typename V>
  ErrNorms(const V err, const V appx, const V exact, const scalar_view_type wt) {
    for (Index i = 0; i < (Index)err.extent(0); ++i) impl::set_err(err, appx, exact, i, false);
    reduce(err, exact, wt);
  }
  template <typename V>
  ErrNorms(const V err, const V appx, const V exact, const scalar_view_type wt, const mask_view_type mask) {
    for (Index i = 0; i < (Index)err.extent(0); ++i) impl::set_err(err, appx, exact, i, mask(i) != 0);
    reduce(err, exact, wt);
  }
  /// for a precomputed error
  template <typename V>
  ErrNorms(const V err, const V exact, const scalar_view_type wt) { reduce(err, exact, wt); }

  std::string info_string(const std::string& label = "", const int tab_level = 0) const {
    std::ostringstream ss;
    ss << std::string(tab_level, '\t') << label << (label.empty() ? "" : " ") << "ErrNorms: l1 = " << std::setprecision(8)
       << l1 << " l2 = " << l2 << " linf = " << linf;
    return ss.str();
  }

 private:
  template <typename V>
  void reduce(const V& err, const V& exact, const scalar_view_type& wt) {
    ENormScalar ll;
    for (Index i = 0; i < (Index)err.extent(0); ++i) {
      const Real e = impl::row_mag(err, i), x = impl::row_mag(exact, i);
      ll.l1num += e * wt(i);
      ll.l1denom += x * wt(i);
      ll.l2num += square(e) * wt(i);
      ll.l2denom += square(x) * wt(i);
      ll.linfnum = (e > ll.linfnum ? e : ll.linfnum);
      ll.linfdenom = (x > ll.linfdenom ? x : ll.linfdenom);
    }
    l1 = ll.l1num / ll.l1denom;
    l2 = std::sqrt(ll.l2num / ll.l2denom);
    linf = ll.linfnum / ll.linfdenom;
  }
};

}  // namespace Lpm
#endif
