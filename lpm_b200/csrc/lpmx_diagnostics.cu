// lpmx_diagnostics.cu -- the callers' per-step O(N) tail on the device: conserved totals of
// the Incompressible2D state and the weighted error norms the example drivers log every step.
//   Incompressible2D::total_vorticity / total_enstrophy / total_kinetic_energy   src/lpm_incompressible2d_impl.hpp:91-137
//   ErrNorms / ReduceErrorFtor                                                  src/lpm_error.hpp:81-131, src/lpm_error_impl.hpp:59-108
// HBM-bound streaming reductions (8-32 bytes per particle); two-pass and deterministic: a fixed grid writes one
// partial per block (shared-memory tree), a single block adds the partials in index order.
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 592;  // 4 x 148

// combine NS sums followed by NM maxima across the block; thread 0 ends with the block's values
template <int NS, int NM>
__device__ __forceinline__ void block_combine(double* v) {
  __shared__ double sh[(NS + NM) * kRedThreads];
#pragma unroll
  for (int q = 0; q < NS + NM; ++q) sh[q * kRedThreads + threadIdx.x] = v[q];
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int q = 0; q < NS; ++q) sh[q * kRedThreads + threadIdx.x] += sh[q * kRedThreads + threadIdx.x + s];
#pragma unroll
      for (int q = NS; q < NS + NM; ++q)
        sh[q * kRedThreads + threadIdx.x] = fmax(sh[q * kRedThreads + threadIdx.x], sh[q * kRedThreads + threadIdx.x + s]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < NS + NM; ++q) v[q] = sh[q * kRedThreads];
}

template <int NS, int NM>
__global__ void __launch_bounds__(kRedThreads) final_combine_kernel(const double* __restrict__ part, int n_blocks,
                                                                    double* __restrict__ out) {
  double v[NS + NM];
#pragma unroll
  for (int q = 0; q < NS + NM; ++q) v[q] = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += kRedThreads) {
#pragma unroll
    for (int q = 0; q < NS; ++q) v[q] += part[(size_t)b * (NS + NM) + q];
#pragma unroll
    for (int q = NS; q < NS + NM; ++q) v[q] = fmax(v[q], part[(size_t)b * (NS + NM) + q]);
  }
  block_combine<NS, NM>(v);
  if (threadIdx.x == 0)
    for (int q = 0; q < NS + NM; ++q) out[q] = v[q];
}

// totals over the leaves: [sum zeta A, sum zeta^2 A, sum |u|^2 A]
__global__ void __launch_bounds__(kRedThreads) ic2d_totals_kernel(int n, const double* __restrict__ zeta, Vec3View u,
                                                                  const double* __restrict__ area,
                                                                  const unsigned char* __restrict__ mask,
                                                                  double* __restrict__ part) {
  double v[3] = {0, 0, 0};
  for (long i = blockIdx.x * (long)kRedThreads + threadIdx.x; i < n; i += (long)gridDim.x * kRedThreads) {
    if (mask[i]) continue;
    const double a = area[i], z = zeta[i];
    v[0] += z * a;
    v[1] += z * z * a;
    v[2] += (u(i, 0) * u(i, 0) + u(i, 1) * u(i, 1) + u(i, 2) * u(i, 2)) * a;
  }
  block_combine<3, 0>(v);
  if (threadIdx.x == 0)
    for (int q = 0; q < 3; ++q) part[(size_t)blockIdx.x * 3 + q] = v[q];
}

// ReduceErrorFtor: [l1num, l1denom, l2num, l2denom | linfnum, linfdenom]
template <int NDIM>
__global__ void __launch_bounds__(kRedThreads) err_norms_kernel(int n, Vec3View err, Vec3View exact,
                                                                const double* __restrict__ weight,
                                                                double* __restrict__ part) {
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (long i = blockIdx.x * (long)kRedThreads + threadIdx.x; i < n; i += (long)gridDim.x * kRedThreads) {
    double e, x;
    if (NDIM == 1) {
      e = fabs(err.p[i]);
      x = fabs(exact.p[i]);
    } else {
      e = sqrt(err(i, 0) * err(i, 0) + err(i, 1) * err(i, 1) + err(i, 2) * err(i, 2));
      x = sqrt(exact(i, 0) * exact(i, 0) + exact(i, 1) * exact(i, 1) + exact(i, 2) * exact(i, 2));
    }
    const double w = weight[i];
    v[0] += e * w;
    v[1] += x * w;
    v[2] += e * e * w;
    v[3] += x * x * w;
    v[4] = fmax(v[4], e);
    v[5] = fmax(v[5], x);
  }
  block_combine<4, 2>(v);
  if (threadIdx.x == 0)
    for (int q = 0; q < 6; ++q) part[(size_t)blockIdx.x * 6 + q] = v[q];
}

static int red_blocks(int n) {
  int b = (n + kRedThreads - 1) / kRedThreads;
  if (b > kRedMaxBlocks) b = kRedMaxBlocks;
  return b < 1 ? 1 : b;
}

int ic2d_totals_device(lpmx_handle_t h, int n, const double* zeta, Vec3View u, const double* area, const unsigned char* mask,
                       double* out3_host) {
  void* scratch = nullptr;
  LPMX_TRY(dev_buffer(h, "red_scratch", sizeof(double) * (6 * (size_t)kRedMaxBlocks + 8), &scratch));
  double* part = (double*)scratch;
  double* res = part + 6 * (size_t)kRedMaxBlocks;
  const int blocks = red_blocks(n);
  ic2d_totals_kernel<<<blocks, kRedThreads, 0, h->stream>>>(n, zeta, u, area, mask, part);
  final_combine_kernel<3, 0><<<1, kRedThreads, 0, h->stream>>>(part, blocks, res);
  h->launches += 2;
  LPMX_CUDA(h, cudaGetLastError());
  LPMX_CUDA(h, cudaMemcpyAsync(out3_host, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

}  // namespace lpmx

extern "C" {

int lpmx_ic2d_totals(lpmx_handle_t h, int n_active, const double* active_vort, const double* active_vel, int layout,
                     long active_ld, const double* active_area, const unsigned char* active_mask, double* total_vorticity,
                     double* total_kinetic_energy, double* total_enstrophy) {
  if (!h) return LPMX_ERR_INVALID;
  if (n_active < 0 || (n_active > 0 && (!active_vort || !active_vel || !active_area || !active_mask)))
    return set_error(h, LPMX_ERR_INVALID, "null array");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (layout == LPMX_LAYOUT_LEFT && active_ld < n_active) return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const size_t vb = (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * active_ld + n_active) : (size_t)3 * n_active) * sizeof(double);
  const void *dz, *du, *da, *dm;
  LPMX_TRY(stage_in(h, "tot_z", active_vort, sizeof(double) * (size_t)n_active, &dz));
  LPMX_TRY(stage_in(h, "tot_u", active_vel, vb, &du));
  LPMX_TRY(stage_in(h, "tot_a", active_area, sizeof(double) * (size_t)n_active, &da));
  LPMX_TRY(stage_in(h, "tot_m", active_mask, (size_t)n_active, &dm));
  double out[3] = {0, 0, 0};
  if (n_active > 0)
    LPMX_TRY(ic2d_totals_device(h, n_active, (const double*)dz, make_view((const double*)du, layout, active_ld),
                                (const double*)da, (const unsigned char*)dm, out));
  if (total_vorticity) *total_vorticity = out[0];
  if (total_enstrophy) *total_enstrophy = 0.5 * out[1];
  if (total_kinetic_energy) *total_kinetic_energy = 0.5 * out[2];
  return LPMX_OK;
}

int lpmx_err_norms(lpmx_handle_t h, int n, int ndim, const double* err, const double* exact, int layout, long ld,
                   const double* weight, double* l1, double* l2, double* linf) {
  if (!h) return LPMX_ERR_INVALID;
  if (n < 0 || (ndim != 1 && ndim != 3) || (n > 0 && (!err || !exact || !weight))) return set_error(h, LPMX_ERR_INVALID, "bad argument");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (ndim == 3 && layout == LPMX_LAYOUT_LEFT && ld < n) return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const size_t vb = ndim == 1 ? sizeof(double) * (size_t)n
                              : (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * ld + n) : (size_t)3 * n) * sizeof(double);
  const void *de, *dx, *dw;
  LPMX_TRY(stage_in(h, "en_err", err, vb, &de));
  LPMX_TRY(stage_in(h, "en_exact", exact, vb, &dx));
  LPMX_TRY(stage_in(h, "en_w", weight, sizeof(double) * (size_t)n, &dw));
  double out[6] = {0, 0, 0, 0, 0, 0};
  if (n > 0) {
    void* scratch = nullptr;
    LPMX_TRY(dev_buffer(h, "red_scratch", sizeof(double) * (6 * (size_t)kRedMaxBlocks + 8), &scratch));
    double* part = (double*)scratch;
    double* res = part + 6 * (size_t)kRedMaxBlocks;
    const int blocks = red_blocks(n);
    const Vec3View ev = make_view((const double*)de, layout, ld), xv = make_view((const double*)dx, layout, ld);
    if (ndim == 1)
      err_norms_kernel<1><<<blocks, kRedThreads, 0, h->stream>>>(n, ev, xv, (const double*)dw, part);
    else
      err_norms_kernel<3><<<blocks, kRedThreads, 0, h->stream>>>(n, ev, xv, (const double*)dw, part);
    final_combine_kernel<4, 2><<<1, kRedThreads, 0, h->stream>>>(part, blocks, res);
    h->launches += 2;
    LPMX_CUDA(h, cudaGetLastError());
    LPMX_CUDA(h, cudaMemcpyAsync(out, res, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // ErrNorms(const ENormScalar&) (src/lpm_error.hpp:96-99): plain ratios, so 0/0 -> NaN exactly as in the reference
  if (l1) *l1 = out[0] / out[1];
  if (l2) *l2 = sqrt(out[2] / out[3]);
  if (linf) *linf = out[4] / out[5];
  return LPMX_OK;
}

}  // extern "C"
