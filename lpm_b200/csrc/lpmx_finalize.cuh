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

// SWE finalize (SphereVertexSums::operator(), lpm_swe_kernels.hpp:756-779 on the factored accumulators of
// Pair<kSwe>): u = x cross Mz + P_x Ms ;  G_total = G + [Mz]x - (x.Ms) P_x ;  ddot = sum_ab G_ab G_ba.
// acc[0..2] = Mz, acc[3..5] = Ms, acc[6..14] = G (row-major).  g9 receives G_total.
__device__ __forceinline__ double swe_finalize(const double* acc, const double* x, double* u, double* g9) {
  const double* mz = acc;
  const double* ms = acc + 3;
  const double xms = x[0] * ms[0] + x[1] * ms[1] + x[2] * ms[2];
  cross3(u, x, mz);
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = u[k] + (ms[k] - xms * x[k]);
  const double mzx[9] = {0, -mz[2], mz[1], mz[2], 0, -mz[0], -mz[1], mz[0], 0};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double P = (a == b ? 1.0 : 0.0) - x[a] * x[b];
      g9[3 * a + b] = acc[6 + 3 * a + b] + mzx[3 * a + b] - xms * P;
    }
  double dd = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) dd += g9[3 * a + b] * g9[3 * b + a];
  return dd;
}

}  // namespace lpmx
#endif
