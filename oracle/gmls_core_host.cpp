// gmls_core_host.cpp -- TEST INFRASTRUCTURE: runs the product's per-target GMLS arithmetic
// (lpm_b200/csrc/lpmx_gmls_core.h, the __host__ __device__ header the CUDA kernel is built from) on the CPU, so the
// `-m "not gpu"` suite can check it against oracle/gmls_oracle.py and the analytic anchors before it ever reaches a
// GPU.  Nothing in the product links or loads this file.  Build: oracle/Makefile -> oracle/_build/libgmls_core_host.so
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../lpm_b200/csrc/lpmx_gmls_core.h"

using namespace lpmx::gmls;

extern "C" int gmls_core_host_laplacian(int n, const double* xyz /* n x 3 row-major */, const double* f, int samples_order,
                                        int manifold_order, int min_neighbors, double eps_multiplier, double weight_pwr,
                                        double radius, double* lap, double* eps_out, int* nn_out) {
  if (n <= 0 || samples_order > kMaxOrder || manifold_order > kMaxOrder || min_neighbors > kMaxK) return -1;
  Params p{samples_order, manifold_order, min_neighbors, eps_multiplier, weight_pwr};
  GridDims gd = grid_dims(n, min_neighbors, eps_multiplier, radius);
  Cloud c;
  c.n = n, c.G = gd.G, c.box = gd.box, c.cell = gd.cell, c.inv_cell = 1.0 / gd.cell;
  std::vector<long> key(n);
  for (int i = 0; i < n; ++i)
    key[i] = ((long)cell_coord(c, xyz[3 * i]) * c.G + cell_coord(c, xyz[3 * i + 1])) * c.G + cell_coord(c, xyz[3 * i + 2]);
  std::vector<int> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<double> xs(3 * (size_t)n), fs(n);
  const long ncell = (long)c.G * c.G * c.G;
  std::vector<int> start(ncell + 1, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) xs[(size_t)k * n + i] = xyz[3 * perm[i] + k];
    fs[i] = f[perm[i]];
    start[key[perm[i]] + 1]++;
  }
  for (long q = 0; q < ncell; ++q) start[q + 1] += start[q];
  c.x = xs.data(), c.f = fs.data(), c.cell_start = start.data();
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    const TargetResult r = laplacian_at_target_dispatch(c, p, i);
    lap[perm[i]] = r.lap;
    if (eps_out) eps_out[perm[i]] = r.eps;
    if (nn_out) nn_out[perm[i]] = r.n_neighbors;
  }
  return 0;
}

// scalar point evaluation of n_fields source fields at n_tgt target points (remeshing); fields are [n_fields][n_src],
// out is [n_fields][n_tgt]
extern "C" int gmls_core_host_interpolate(int n_src, const double* src_xyz, int n_fields, const double* src_fields, int n_tgt,
                                          const double* tgt_xyz, int samples_order, int min_neighbors, double eps_multiplier,
                                          double weight_pwr, double radius, double* out) {
  if (n_src <= 0 || samples_order < 1 || samples_order > kMaxOrder || min_neighbors > kMaxK) return -1;
  Params p{samples_order, samples_order, min_neighbors, eps_multiplier, weight_pwr};
  GridDims gd = grid_dims(n_src, min_neighbors, eps_multiplier, radius);
  Cloud c;
  c.n = n_src, c.G = gd.G, c.box = gd.box, c.cell = gd.cell, c.inv_cell = 1.0 / gd.cell;
  std::vector<long> key(n_src);
  for (int i = 0; i < n_src; ++i)
    key[i] = ((long)cell_coord(c, src_xyz[3 * i]) * c.G + cell_coord(c, src_xyz[3 * i + 1])) * c.G + cell_coord(c, src_xyz[3 * i + 2]);
  std::vector<int> perm(n_src);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<double> xs(3 * (size_t)n_src);
  const long ncell = (long)c.G * c.G * c.G;
  std::vector<int> start(ncell + 1, 0);
  for (int i = 0; i < n_src; ++i) {
    for (int k = 0; k < 3; ++k) xs[(size_t)k * n_src + i] = src_xyz[3 * perm[i] + k];
    start[key[perm[i]] + 1]++;
  }
  for (long q = 0; q < ncell; ++q) start[q + 1] += start[q];
  c.x = xs.data(), c.f = nullptr, c.cell_start = start.data();
  std::vector<double> zeros(n_src, 0.0);
  for (int f0 = 0; f0 < n_fields; f0 += kInterpFields) {
    std::vector<std::vector<double>> sorted(kInterpFields);
    Fields fl;
    for (int q = 0; q < kInterpFields; ++q) {
      if (f0 + q < n_fields) {
        sorted[q].resize(n_src);
        for (int i = 0; i < n_src; ++i) sorted[q][i] = src_fields[(size_t)(f0 + q) * n_src + perm[i]];
        fl.f[q] = sorted[q].data();
      } else {
        fl.f[q] = zeros.data();
      }
    }
    const int om = samples_order < 2 ? 2 : samples_order;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n_tgt; ++i) {
      double v[kInterpFields];
      const double *x = tgt_xyz + 3 * (size_t)i;
      if (om == 2) interpolate_at_point<2, kMaxK>(c, fl, p, x[0], x[1], x[2], v);
      else if (om == 3) interpolate_at_point<3, kMaxK>(c, fl, p, x[0], x[1], x[2], v);
      else interpolate_at_point<4, kMaxK>(c, fl, p, x[0], x[1], x[2], v);
      for (int q = 0; q < kInterpFields && f0 + q < n_fields; ++q) out[(size_t)(f0 + q) * n_tgt + i] = v[q];
    }
  }
  return 0;
}
