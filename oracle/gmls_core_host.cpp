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
