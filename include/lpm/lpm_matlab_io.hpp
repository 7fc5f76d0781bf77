// lpm/lpm_matlab_io.hpp -- write_vector_matlab / write_array_matlab (src/util/lpm_matlab_io.hpp:11-52).
// Same text as the reference: `name = [v0,v1,...];\n` and `name = [a00,a01;a10,a11];\n`, numbers through
// operator<<(double) at the stream's current precision (6 significant digits unless the caller changes it).
// Pinned byte-for-byte against the reference header compiled in place (tests/test_io_formats.py).
#ifndef LPM_SHIM_MATLAB_IO_HPP
#define LPM_SHIM_MATLAB_IO_HPP

#include <iostream>
#include <string>
#include <vector>

#include "lpm_views.hpp"

namespace Lpm {

template <typename HVT>
inline void write_vector_matlab(std::ostream& os, const std::string& name, const HVT& v) {
  const auto last_idx = v.extent(0) - 1;
  os << name << " = [";
  for (size_t i = 0; i < last_idx; ++i) os << v(i) << ",";
  os << v(last_idx) << "];\n";
}

template <>
inline void write_vector_matlab<std::vector<Real>>(std::ostream& os, const std::string& name, const std::vector<Real>& v) {
  const auto last_idx = v.size() - 1;
  os << name << " = [";
  for (size_t i = 0; i < last_idx; ++i) os << v[i] << ",";
  os << v[last_idx] << "];\n";
}

template <typename HVT>
inline void write_array_matlab(std::ostream& os, const std::string name, const HVT a) {
  const auto nrow = a.extent(0), ncol = a.extent(1);
  const auto last_row = nrow - 1, last_col = ncol - 1;
  os << name << " = [";
  for (size_t i = 0; i < nrow; ++i)
    for (size_t j = 0; j < ncol; ++j)
      os << a(i, j) << (i < last_row ? (j < last_col ? "," : ";") : (j < last_col ? "," : "];\n"));
}

}  // namespace Lpm
#endif
