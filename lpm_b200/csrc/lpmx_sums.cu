// lpmx_sums.cu -- operator-level (stateless) direct sums behind the C ABI.
//
// Each entry point = pack sources (leaf compaction + Gamma) -> pair-sum kernel -> finalize.
// Reference functors replaced: see include/lpmx.h.
#include <cfloat>
#include <cmath>

#include "lpmx_finalize.cuh"
#include "lpmx_internal.h"

namespace lpmx {

// ------------------------------------------------------------------------------------------------
// pack: faces -> packed leaf-only source records; optional self index per (collocated) target
// ------------------------------------------------------------------------------------------------
template <int REC>
__global__ void pack_sources_kernel(Vec3View sx, const double* __restrict__ vort, const double* __restrict__ div,
                                    const double* __restrict__ area, const unsigned char* __restrict__ mask,
                                    const int* __restrict__ leaf_idx, int n_src, int n_leaf, int n_src_pad,
                                    double* __restrict__ packed, int* __restrict__ self_idx, int skip_self) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += stride) {
    const bool leaf = mask[i] == 0;
    if (leaf) {
      double* rec = packed + (size_t)leaf_idx[i] * REC;
      const double y[3] = {sx(i, 0), sx(i, 1), sx(i, 2)};
      if (REC == 6) {
        rec[0] = y[0];
        rec[1] = y[1];
        rec[2] = y[2];
        rec[3] = gamma_of(vort[i], area[i]);
        rec[4] = gamma_of(div[i], area[i]);
        rec[5] = 0.0;
      } else {
        write_bve_record(rec, y, gamma_of(vort[i], area[i]));
      }
    }
    if (self_idx) self_idx[i] = (skip_self && leaf) ? leaf_idx[i] : -1;
  }
  // zero-strength padding at the origin: d = kappa, w = 0 -> contributes exactly 0
  for (int i = n_leaf + blockIdx.x * blockDim.x + threadIdx.x; i < n_src_pad; i += stride) {
    double* rec = packed + (size_t)i * REC;
#pragma unroll
    for (int k = 0; k < REC; ++k) rec[k] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// finalize kernels for the stateless entry points
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void fin_plain_kernel(PartView pv, int n_tgt, Vec3View tx, Vec3View out_vel, double* __restrict__ out_psi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tgt) return;
  constexpr int NACC = kind_nacc(KIND);
  double acc[NACC];
  reduce_slots<NACC>(pv, i, acc);
  if (KIND == kVel || KIND == kVelPsi) {
    const double x[3] = {tx(i, 0), tx(i, 1), tx(i, 2)};
    double u[3];
    cross3(u, x, acc);
    if (out_vel.p) {
      out_vel(i, 0) = u[0];
      out_vel(i, 1) = u[1];
      out_vel(i, 2) = u[2];
    }
    if (KIND == kVelPsi && out_psi) out_psi[i] = acc[3];
  } else if (KIND == kPsi) {
    out_psi[i] = acc[0];
  }
}

// SWE: u = x cross Mz + P_x Ms ;  G_total = G + [Mz]x - (x.Ms) P_x ;  ddot = sum_ab G_ab G_ba
// (SphereVertexSums::operator(), lpm_swe_kernels.hpp:756-779)
__global__ void fin_swe_kernel(PartView pv, int n_tgt, Vec3View tx, int do_velocity, Vec3View out_vel,
                               double* __restrict__ out_ddot, double* __restrict__ out_grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tgt) return;
  double acc[15];
  reduce_slots<15>(pv, i, acc);
  const double x[3] = {tx(i, 0), tx(i, 1), tx(i, 2)};
  double u[3], g[9];
  const double dd = swe_finalize(acc, x, u, g);
  if (do_velocity && out_vel.p)
    for (int k = 0; k < 3; ++k) out_vel(i, k) = u[k];
  out_ddot[i] = dd;
  if (out_grad)
    for (int k = 0; k < 9; ++k) out_grad[9L * i + k] = g[k];
}

static size_t vec_bytes(int layout, long ld, int n) {
  return (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * ld + n) : (size_t)3 * n) * sizeof(double);
}

struct SumCall {
  int kind;
  const double* tgt_xyz;
  int tgt_layout;
  long tgt_ld;
  int n_tgt;
  const double* src_xyz;
  int src_layout;
  long src_ld;
  const double* src_vort;
  const double* src_div;
  const double* src_area;
  const unsigned char* src_mask;
  int n_src;
  double eps;
  int targets_are_sources;
  int skip_self;
  int do_velocity;
  double* out_vel;
  double* out_psi;   // psi or ddot
  double* out_grad;  // swe only
};

static int run_sum(lpmx_handle_t h, const SumCall& c) {
  if (!h) return LPMX_ERR_INVALID;
  if (c.n_src < 0 || c.n_tgt < 0) return set_error(h, LPMX_ERR_INVALID, "negative size");
  if (c.n_src > 0 && (!c.src_xyz || !c.src_vort || !c.src_area || !c.src_mask))
    return set_error(h, LPMX_ERR_INVALID, "null source array");
  if (c.kind == kSwe && c.n_src > 0 && !c.src_div) return set_error(h, LPMX_ERR_INVALID, "null divergence array");
  if (c.targets_are_sources && c.n_tgt != c.n_src)
    return set_error(h, LPMX_ERR_INVALID, "collocated call needs n_tgt == n_src");
  if (!c.targets_are_sources && c.n_tgt > 0 && !c.tgt_xyz) return set_error(h, LPMX_ERR_INVALID, "null target array");
  if ((c.tgt_layout != LPMX_LAYOUT_LEFT && c.tgt_layout != LPMX_LAYOUT_RIGHT) ||
      (c.src_layout != LPMX_LAYOUT_LEFT && c.src_layout != LPMX_LAYOUT_RIGHT))
    return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((c.src_layout == LPMX_LAYOUT_LEFT && c.src_ld < c.n_src) ||
      (!c.targets_are_sources && c.tgt_layout == LPMX_LAYOUT_LEFT && c.tgt_ld < c.n_tgt))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  if (c.n_tgt == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));

  const int t_layout = c.targets_are_sources ? c.src_layout : c.tgt_layout;
  const long t_ld = c.targets_are_sources ? c.src_ld : c.tgt_ld;
  bool any_host = false;
  auto host = [&](const void* p) { return p && !is_device_pointer(p); };
  any_host = host(c.src_xyz) || host(c.src_vort) || host(c.src_area) || host(c.src_mask) || host(c.src_div) ||
             host(c.tgt_xyz) || host(c.out_vel) || host(c.out_psi) || host(c.out_grad);

  const void *d_sx = nullptr, *d_vort = nullptr, *d_div = nullptr, *d_area = nullptr, *d_mask = nullptr, *d_tx = nullptr;
  LPMX_TRY(stage_in(h, "in_src_xyz", c.src_xyz, vec_bytes(c.src_layout, c.src_ld, c.n_src), &d_sx));
  LPMX_TRY(stage_in(h, "in_src_vort", c.src_vort, sizeof(double) * c.n_src, &d_vort));
  LPMX_TRY(stage_in(h, "in_src_area", c.src_area, sizeof(double) * c.n_src, &d_area));
  LPMX_TRY(stage_in(h, "in_src_mask", c.src_mask, (size_t)c.n_src, &d_mask));
  if (c.kind == kSwe) LPMX_TRY(stage_in(h, "in_src_div", c.src_div, sizeof(double) * c.n_src, &d_div));
  if (c.targets_are_sources)
    d_tx = d_sx;
  else
    LPMX_TRY(stage_in(h, "in_tgt_xyz", c.tgt_xyz, vec_bytes(c.tgt_layout, c.tgt_ld, c.n_tgt), &d_tx));

  // leaf compaction
  void* d_leaf = nullptr;
  LPMX_TRY(dev_buffer(h, "leaf_idx", sizeof(int) * (size_t)(c.n_src + 1), &d_leaf));
  int n_leaf = 0;
  LPMX_TRY(scan_leaves(h, (const unsigned char*)d_mask, c.n_src, (int*)d_leaf, &n_leaf));

  SumPlan plan;
  LPMX_TRY(make_plan(h, c.kind, c.n_tgt, n_leaf, &plan));
  const int rec = kind_rec(c.kind);
  void *d_packed = nullptr, *d_self = nullptr, *d_part = nullptr;
  LPMX_TRY(dev_buffer(h, "packed", sizeof(double) * rec * (size_t)(plan.n_src_pad + kChunk), &d_packed));
  LPMX_TRY(dev_buffer(h, "partials", plan_partials_bytes(plan) + 256, &d_part));
  if (c.skip_self) LPMX_TRY(dev_buffer(h, "self_idx", sizeof(int) * (size_t)(c.n_src + 1), &d_self));

  const Vec3View sxv = make_view((const double*)d_sx, c.src_layout, c.src_ld);
  const Vec3View txv = make_view((const double*)d_tx, t_layout, t_ld);
  {
    const int threads = 256;
    int blocks = (std::max(c.n_src, plan.n_src_pad - n_leaf) + threads - 1) / threads;
    if (blocks < 1) blocks = 1;
    if (rec == kBveRec)
      pack_sources_kernel<kBveRec><<<blocks, threads, 0, h->stream>>>(sxv, (const double*)d_vort, nullptr, (const double*)d_area,
                                                                (const unsigned char*)d_mask, (const int*)d_leaf,
                                                                c.n_src, n_leaf, plan.n_src_pad, (double*)d_packed,
                                                                (int*)d_self, c.skip_self);
    else
      pack_sources_kernel<6><<<blocks, threads, 0, h->stream>>>(sxv, (const double*)d_vort, (const double*)d_div,
                                                                (const double*)d_area, (const unsigned char*)d_mask,
                                                                (const int*)d_leaf, c.n_src, n_leaf, plan.n_src_pad,
                                                                (double*)d_packed, (int*)d_self, c.skip_self);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  const double kappa = 1.0 + c.eps * c.eps;
  LPMX_TRY(launch_pair_sum(h, plan, txv, (const int*)d_self, (const double*)d_packed, kappa, (double*)d_part));

  // outputs
  void *d_vel = nullptr, *d_psi = nullptr, *d_grad = nullptr;
  const size_t vel_bytes = vec_bytes(t_layout, t_ld, c.n_tgt);
  if (c.out_vel) LPMX_TRY(stage_out_begin(h, "out_vel", c.out_vel, vel_bytes, &d_vel));
  if (c.out_psi) LPMX_TRY(stage_out_begin(h, "out_psi", c.out_psi, sizeof(double) * c.n_tgt, &d_psi));
  if (c.out_grad) LPMX_TRY(stage_out_begin(h, "out_grad", c.out_grad, sizeof(double) * 9 * (size_t)c.n_tgt, &d_grad));
  const PartView pv = part_view(plan, (const double*)d_part);
  const Vec3View ov = make_view((const double*)d_vel, t_layout, t_ld);
  {
    const int threads = 128;
    const int blocks = (c.n_tgt + threads - 1) / threads;
    switch (c.kind) {
      case kVel: fin_plain_kernel<kVel><<<blocks, threads, 0, h->stream>>>(pv, c.n_tgt, txv, ov, (double*)d_psi); break;
      case kVelPsi:
        fin_plain_kernel<kVelPsi><<<blocks, threads, 0, h->stream>>>(pv, c.n_tgt, txv, ov, (double*)d_psi);
        break;
      case kPsi: fin_plain_kernel<kPsi><<<blocks, threads, 0, h->stream>>>(pv, c.n_tgt, txv, ov, (double*)d_psi); break;
      case kSwe:
        fin_swe_kernel<<<blocks, threads, 0, h->stream>>>(pv, c.n_tgt, txv, c.do_velocity, ov, (double*)d_psi,
                                                          (double*)d_grad);
        break;
    }
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  if (c.out_vel && (c.kind != kSwe || c.do_velocity)) LPMX_TRY(stage_out_end(h, c.out_vel, d_vel, vel_bytes));
  if (c.out_psi) LPMX_TRY(stage_out_end(h, c.out_psi, d_psi, sizeof(double) * c.n_tgt));
  if (c.out_grad) LPMX_TRY(stage_out_end(h, c.out_grad, d_grad, sizeof(double) * 9 * (size_t)c.n_tgt));
  if (any_host) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

}  // namespace lpmx

using namespace lpmx;

extern "C" {

int lpmx_bve_velocity(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                      const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                      const double* src_area, const unsigned char* src_mask, int n_src, int collocated,
                      double* out_vel) {
  if (!h) return LPMX_ERR_INVALID;
  if (!out_vel && n_tgt > 0) return set_error(h, LPMX_ERR_INVALID, "null output");
  SumCall c{kVel,     tgt_xyz,  tgt_layout, tgt_ld,   n_tgt, src_xyz,         src_layout,      src_ld, src_vort,
            nullptr,  src_area, src_mask,   n_src,    0.0,   collocated != 0, collocated != 0, 1,      out_vel,
            nullptr,  nullptr};
  return run_sum(h, c);
}

int lpmx_bve_streamfn(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                      const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                      const double* src_area, const unsigned char* src_mask, int n_src, int collocated,
                      double* out_psi) {
  if (!h) return LPMX_ERR_INVALID;
  if (!out_psi && n_tgt > 0) return set_error(h, LPMX_ERR_INVALID, "null output");
  SumCall c{kPsi,    tgt_xyz,  tgt_layout, tgt_ld, n_tgt, src_xyz,         src_layout,      src_ld, src_vort,
            nullptr, src_area, src_mask,   n_src,  0.0,   collocated != 0, collocated != 0, 0,      nullptr,
            out_psi, nullptr};
  return run_sum(h, c);
}

int lpmx_bve_solve(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt, const double* src_xyz,
                   int src_layout, long src_ld, const double* src_vort, const double* src_area,
                   const unsigned char* src_mask, int n_src, int collocated, double* out_psi, double* out_vel) {
  if (!h) return LPMX_ERR_INVALID;
  if ((!out_vel || !out_psi) && n_tgt > 0) return set_error(h, LPMX_ERR_INVALID, "null output");
  SumCall c{kVelPsi, tgt_xyz,  tgt_layout, tgt_ld, n_tgt, src_xyz,         src_layout,      src_ld, src_vort,
            nullptr, src_area, src_mask,   n_src,  0.0,   collocated != 0, collocated != 0, 1,      out_vel,
            out_psi, nullptr};
  return run_sum(h, c);
}

int lpmx_ic2d_sums(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                   const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                   const double* src_area, const unsigned char* src_mask, int n_src, double eps,
                   int targets_are_sources, double* out_vel, double* out_psi) {
  if (!h) return LPMX_ERR_INVALID;
  if (!out_vel && n_tgt > 0) return set_error(h, LPMX_ERR_INVALID, "null output");
  // Incompressible2DActiveSums: collocated = FloatingPoint<Real>::zero(eps) (:235)
  const int skip = targets_are_sources && (std::fabs(eps) < DBL_EPSILON);
  SumCall c{out_psi ? kVelPsi : kVel, tgt_xyz, tgt_layout, tgt_ld, n_tgt, src_xyz, src_layout, src_ld, src_vort,
            nullptr, src_area, src_mask, n_src, eps, targets_are_sources != 0, skip, 1, out_vel, out_psi, nullptr};
  return run_sum(h, c);
}

int lpmx_swe_sphere_sums(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                         const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                         const double* src_div, const double* src_area, const unsigned char* src_mask, int n_src,
                         double eps, int targets_are_sources, int do_velocity, double* out_vel, double* out_ddot,
                         double* out_grad) {
  if (!h) return LPMX_ERR_INVALID;
  if (n_tgt > 0 && (!out_ddot || (do_velocity && !out_vel))) return set_error(h, LPMX_ERR_INVALID, "null output");
  const int skip = targets_are_sources && (std::fabs(eps) < DBL_EPSILON);
  SumCall c{kSwe,    tgt_xyz,  tgt_layout, tgt_ld, n_tgt, src_xyz, src_layout, src_ld, src_vort,
            src_div, src_area, src_mask,   n_src,  eps,   targets_are_sources != 0, skip, do_velocity,
            do_velocity ? out_vel : nullptr, out_ddot, out_grad};
  return run_sum(h, c);
}

}  // extern "C"
