// lpmx_finalize.cuh -- device helpers shared by the O(N) finalize / pack kernels.
#ifndef LPMX_FINALIZE_CUH
#define LPMX_FINALIZE_CUH

#include "lpmx_internal.h"

namespace lpmx {

// 1/(4 pi) with the reference's PI literal (lpm_constants.hpp:11)
#define LPMX_PI 3.1415926535897932384626433832795027975
__device__ __forceinline__ double gamma_of(double strength, double area) { return (-strength * area) / (4.0 * LPMX_PI); }

// the 64-byte source record of the BVE / IC2D kinds
__device__ __forceinline__ void write_bve_record(double* rec, const double* y, double gam) {
  double2* r2 = reinterpret_cast<double2*>(rec);
  r2[0] = make_double2(y[0], y[1]);
  r2[1] = make_double2(y[2], gam * y[0]);
  r2[2] = make_double2(gam * y[1], gam * y[2]);
  r2[3] = make_double2(gam, 0.0);
}

// Where the pair-sum kernel left its partial sums for a launch (mirrors SumPlan).
struct PartView {
  const double* part;
  long n_tgt_pad;
  long n_items;
  int n_sc;
  int tb;
  int grid;
};
inline PartView part_view(const SumPlan& p, const double* partials) {
  PartView v;
  v.part = partials;
  v.n_tgt_pad = p.n_tgt_pad;
  v.n_items = (long)p.n_tb * p.n_sc;
  v.n_sc = p.n_sc;
  v.tb = p.tb;
  v.grid = p.grid;
  return v;
}

__device__ __forceinline__ int fin_cta_of_item(long item, int grid, long n_items) {
  return (int)(((item + 1) * (long)grid - 1) / n_items);
}

// Sum the slots of target `tg` (launch-local index) in slot order: deterministic.
template <int NACC>
__device__ __forceinline__ void reduce_slots(const PartView& v, long tg, double* acc) {
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
  if (v.n_items == 0) return;
  const int tb = (int)(tg / v.tb);
  const int c0 = fin_cta_of_item((long)tb * v.n_sc, v.grid, v.n_items);
  const int c1 = fin_cta_of_item((long)(tb + 1) * v.n_sc - 1, v.grid, v.n_items);
  for (int slot = 0; slot <= c1 - c0; ++slot) {
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] += v.part[((long)slot * NACC + q) * v.n_tgt_pad + tg];
  }
}

__device__ __forceinline__ void cross3(double* c, const double* a, const double* b) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

}  // namespace lpmx
#endif
