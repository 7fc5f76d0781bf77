// lpmx_refinement.cu -- adaptive-refinement flag functors on the device (the step between the direct sums and
// PolyMesh2d::divide_flagged_faces in the AMR drivers, examples/sphere_gaussian_vortex.cpp:89-118, sphere_rh54.cpp:118-147).
//   ScalarMaxFlag / ScalarIntegralFlag / ScalarVariationFlag / FlowMapVariationFlag   src/mesh/lpm_refinement_flags.hpp:55-310
//   Refinement<Seed>::iterate                                                        src/mesh/lpm_refinement.hpp:28-41
// One thread per face; HBM-bound streaming kernels (10-18 bytes per face for the scalar kinds, plus 3-4 vertex gathers that
// hit L2 for the variation kinds).  The maximum behind set_tol_from_relative_value() is a two-pass reduction (one partial per
// CTA, one CTA combines them in index order); the flag count is an integer atomic, so both are deterministic.
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

constexpr int kFlagThreads = 256;
constexpr int kFlagMaxBlocks = 592;  // 4 x 148
constexpr double kLowest = -1.7976931348623157e308;  // Kokkos::reduction_identity<double>::max()

struct FlagArgs {
  int n_faces, nfv, ndim;
  const double* face_vals;
  const double* area;
  const double* vert_vals;
  const int* face_verts;
  const double* lag;
  long lag_si, lag_sk;
  const unsigned char* mask;
};

// the quantity each functor compares with tol; *active = whether face i takes part in the maximum
template <int KIND>
__device__ __forceinline__ double flag_value(const FlagArgs& a, long i, bool masked) {
  if (KIND == LPMX_FLAG_SCALAR_MAX) return fabs(a.face_vals[i]);
  if (KIND == LPMX_FLAG_SCALAR_INTEGRAL) return fabs(a.face_vals[i]) * a.area[i];
  if (masked) return 0.0;  // the variation kinds never read a divided face
  if (KIND == LPMX_FLAG_SCALAR_VARIATION) {
    double lo = a.face_vals[i], hi = lo;
    for (int j = 0; j < a.nfv; ++j) {
      const double v = a.vert_vals[a.face_verts[i * a.nfv + j]];
      if (v < lo) lo = v;
      if (v > hi) hi = v;
    }
    return hi - lo;
  }
  // FlowMapVariationFlag: extent of the face's vertices in each Lagrangian coordinate, summed over the coordinates
  double lo[3], hi[3];
  const long v0 = a.face_verts[i * a.nfv];
#pragma unroll
  for (int k = 0; k < 3; ++k) lo[k] = hi[k] = k < a.ndim ? a.lag[v0 * a.lag_si + k * a.lag_sk] : 0.0;
  for (int j = 1; j < a.nfv; ++j) {
    const long v = a.face_verts[i * a.nfv + j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < a.ndim) {
        const double x = a.lag[v * a.lag_si + k * a.lag_sk];
        if (x < lo[k]) lo[k] = x;
        if (x > hi[k]) hi[k] = x;
      }
    }
  }
  double dsum = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < a.ndim) dsum += hi[k] - lo[k];
  return dsum;
}

template <int KIND>
__global__ void __launch_bounds__(kFlagThreads) flag_max_kernel(FlagArgs a, double* __restrict__ part) {
  __shared__ double sh[kFlagThreads];
  double m = kLowest;
  for (long i = blockIdx.x * (long)kFlagThreads + threadIdx.x; i < a.n_faces; i += (long)gridDim.x * kFlagThreads) {
    const bool masked = a.mask[i] != 0;
    // the two pointwise kinds reduce over every face, divided ones included (:163-170, :208-216)
    if (masked && (KIND == LPMX_FLAG_SCALAR_VARIATION || KIND == LPMX_FLAG_FLOW_MAP_VARIATION)) continue;
    const double v = flag_value<KIND>(a, i, masked);
    m = v > m ? v : m;
  }
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = kFlagThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x + s] > sh[threadIdx.x] ? sh[threadIdx.x + s] : sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(kFlagThreads) flag_max_final_kernel(const double* __restrict__ part, int n_blocks,
                                                                      double* __restrict__ out) {
  __shared__ double sh[kFlagThreads];
  double m = kLowest;
  for (int b = threadIdx.x; b < n_blocks; b += kFlagThreads) m = part[b] > m ? part[b] : m;
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = kFlagThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x + s] > sh[threadIdx.x] ? sh[threadIdx.x + s] : sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

template <int KIND>
__global__ void __launch_bounds__(kFlagThreads) flag_kernel(FlagArgs a, double tol, int start, int end,
                                                            unsigned char* __restrict__ flags, int* __restrict__ count) {
  int local = 0;
  for (long i = start + blockIdx.x * (long)kFlagThreads + threadIdx.x; i < end; i += (long)gridDim.x * kFlagThreads) {
    unsigned char f = flags[i];
    if (!a.mask[i] && flag_value<KIND>(a, i, false) > tol) f = 1;
    flags[i] = f;
    local += f ? 1 : 0;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

static int flag_blocks(long n) {
  long b = (n + kFlagThreads - 1) / kFlagThreads;
  if (b > kFlagMaxBlocks) b = kFlagMaxBlocks;
  return b < 1 ? 1 : (int)b;
}

// validate the descriptor and stage its arrays; fills `a`
static int stage_flag(lpmx_handle_t h, const lpmx_flag_desc_t* d, FlagArgs* a) {
  if (!d) return set_error(h, LPMX_ERR_INVALID, "null flag descriptor");
  if (d->kind < LPMX_FLAG_SCALAR_MAX || d->kind > LPMX_FLAG_FLOW_MAP_VARIATION) return set_error(h, LPMX_ERR_INVALID, "unknown flag kind");
  if (d->n_faces < 0 || d->n_verts < 0) return set_error(h, LPMX_ERR_INVALID, "negative extent");
  const bool needs_verts = d->kind == LPMX_FLAG_SCALAR_VARIATION || d->kind == LPMX_FLAG_FLOW_MAP_VARIATION;
  if (needs_verts && d->n_face_verts != 3 && d->n_face_verts != 4) return set_error(h, LPMX_ERR_INVALID, "faces have 3 or 4 vertices");
  if (d->n_faces > 0) {
    if (!d->mask) return set_error(h, LPMX_ERR_INVALID, "null mask");
    if (d->kind != LPMX_FLAG_FLOW_MAP_VARIATION && !d->face_vals) return set_error(h, LPMX_ERR_INVALID, "null face values");
    if (d->kind == LPMX_FLAG_SCALAR_INTEGRAL && !d->area) return set_error(h, LPMX_ERR_INVALID, "null area");
    if (d->kind == LPMX_FLAG_SCALAR_VARIATION && !d->vert_vals) return set_error(h, LPMX_ERR_INVALID, "null vertex values");
    if (needs_verts && !d->face_verts) return set_error(h, LPMX_ERR_INVALID, "null face_verts");
    if (d->kind == LPMX_FLAG_FLOW_MAP_VARIATION) {
      if (!d->vert_lag) return set_error(h, LPMX_ERR_INVALID, "null Lagrangian coordinates");
      if (d->ndim != 2 && d->ndim != 3) return set_error(h, LPMX_ERR_INVALID, "ndim is 2 or 3");
      if (d->layout != LPMX_LAYOUT_LEFT && d->layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
      if (d->layout == LPMX_LAYOUT_LEFT && d->ld < d->n_verts) return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
    }
  }
  LPMX_CUDA(h, cudaSetDevice(h->device));
  *a = FlagArgs{};
  a->n_faces = d->n_faces, a->nfv = d->n_face_verts, a->ndim = d->ndim;
  const size_t nf = (size_t)d->n_faces, nv = (size_t)d->n_verts;
  const void* p;
  LPMX_TRY(stage_in(h, "flag_mask", d->mask, nf, &p));
  a->mask = (const unsigned char*)p;
  if (d->kind != LPMX_FLAG_FLOW_MAP_VARIATION) {
    LPMX_TRY(stage_in(h, "flag_fvals", d->face_vals, sizeof(double) * nf, &p));
    a->face_vals = (const double*)p;
  }
  if (d->kind == LPMX_FLAG_SCALAR_INTEGRAL) {
    LPMX_TRY(stage_in(h, "flag_area", d->area, sizeof(double) * nf, &p));
    a->area = (const double*)p;
  }
  if (d->kind == LPMX_FLAG_SCALAR_VARIATION) {
    LPMX_TRY(stage_in(h, "flag_vvals", d->vert_vals, sizeof(double) * nv, &p));
    a->vert_vals = (const double*)p;
  }
  if (needs_verts) {
    LPMX_TRY(stage_in(h, "flag_fverts", d->face_verts, sizeof(int) * nf * (size_t)d->n_face_verts, &p));
    a->face_verts = (const int*)p;
  }
  if (d->kind == LPMX_FLAG_FLOW_MAP_VARIATION) {
    const size_t bytes = (d->layout == LPMX_LAYOUT_LEFT ? (size_t)((d->ndim - 1) * d->ld) + nv : (size_t)d->ndim * nv) * sizeof(double);
    LPMX_TRY(stage_in(h, "flag_lag", d->vert_lag, bytes, &p));
    a->lag = (const double*)p;
    if (d->layout == LPMX_LAYOUT_LEFT) a->lag_si = 1, a->lag_sk = d->ld;
    else a->lag_si = d->ndim, a->lag_sk = 1;
  }
  return LPMX_OK;
}

}  // namespace lpmx

extern "C" {

int lpmx_refine_flag_max(lpmx_handle_t h, const lpmx_flag_desc_t* d, double* max_value) {
  if (!h) return LPMX_ERR_INVALID;
  if (!max_value) return set_error(h, LPMX_ERR_INVALID, "null result");
  FlagArgs a;
  LPMX_TRY(stage_flag(h, d, &a));
  double mx = kLowest;
  if (d->n_faces > 0) {
    void* scratch = nullptr;
    LPMX_TRY(dev_buffer(h, "flag_scratch", sizeof(double) * ((size_t)kFlagMaxBlocks + 8), &scratch));
    double* part = (double*)scratch;
    double* res = part + kFlagMaxBlocks;
    const int blocks = flag_blocks(d->n_faces);
    switch (d->kind) {
      case LPMX_FLAG_SCALAR_MAX: flag_max_kernel<LPMX_FLAG_SCALAR_MAX><<<blocks, kFlagThreads, 0, h->stream>>>(a, part); break;
      case LPMX_FLAG_SCALAR_INTEGRAL: flag_max_kernel<LPMX_FLAG_SCALAR_INTEGRAL><<<blocks, kFlagThreads, 0, h->stream>>>(a, part); break;
      case LPMX_FLAG_SCALAR_VARIATION: flag_max_kernel<LPMX_FLAG_SCALAR_VARIATION><<<blocks, kFlagThreads, 0, h->stream>>>(a, part); break;
      default: flag_max_kernel<LPMX_FLAG_FLOW_MAP_VARIATION><<<blocks, kFlagThreads, 0, h->stream>>>(a, part); break;
    }
    flag_max_final_kernel<<<1, kFlagThreads, 0, h->stream>>>(part, blocks, res);
    h->launches += 2;
    LPMX_CUDA(h, cudaGetLastError());
    LPMX_CUDA(h, cudaMemcpyAsync(&mx, res, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  *max_value = mx;
  return LPMX_OK;
}

int lpmx_refine_flag(lpmx_handle_t h, const lpmx_flag_desc_t* d, int start, int end, int clear_first, unsigned char* flags,
                     int* count) {
  if (!h) return LPMX_ERR_INVALID;
  FlagArgs a;
  LPMX_TRY(stage_flag(h, d, &a));
  if (start < 0 || end < start || end > d->n_faces) return set_error(h, LPMX_ERR_INVALID, "face range outside [0, n_faces]");
  if (d->n_faces > 0 && !flags) return set_error(h, LPMX_ERR_INVALID, "null flags");
  int ct = 0;
  if (d->n_faces > 0) {
    const size_t nf = (size_t)d->n_faces;
    // flags are in/out: staged in, copied back when they live on the host
    const void* tmp;
    LPMX_TRY(stage_in(h, "flag_flags", flags, nf, &tmp));
    unsigned char* dflags = (unsigned char*)const_cast<void*>(tmp);
    void* scratch = nullptr;
    LPMX_TRY(dev_buffer(h, "flag_count", 16, &scratch));
    int* dcount = (int*)scratch;
    if (clear_first) LPMX_CUDA(h, cudaMemsetAsync(dflags, 0, nf, h->stream));
    LPMX_CUDA(h, cudaMemsetAsync(dcount, 0, sizeof(int), h->stream));
    if (end > start) {
      const int blocks = flag_blocks(end - start);
      switch (d->kind) {
        case LPMX_FLAG_SCALAR_MAX: flag_kernel<LPMX_FLAG_SCALAR_MAX><<<blocks, kFlagThreads, 0, h->stream>>>(a, d->tol, start, end, dflags, dcount); break;
        case LPMX_FLAG_SCALAR_INTEGRAL: flag_kernel<LPMX_FLAG_SCALAR_INTEGRAL><<<blocks, kFlagThreads, 0, h->stream>>>(a, d->tol, start, end, dflags, dcount); break;
        case LPMX_FLAG_SCALAR_VARIATION: flag_kernel<LPMX_FLAG_SCALAR_VARIATION><<<blocks, kFlagThreads, 0, h->stream>>>(a, d->tol, start, end, dflags, dcount); break;
        default: flag_kernel<LPMX_FLAG_FLOW_MAP_VARIATION><<<blocks, kFlagThreads, 0, h->stream>>>(a, d->tol, start, end, dflags, dcount); break;
      }
      h->launches += 1;
      LPMX_CUDA(h, cudaGetLastError());
    }
    if (dflags != flags) LPMX_TRY(stage_out_end(h, flags, dflags, nf));
    LPMX_CUDA(h, cudaMemcpyAsync(&ct, dcount, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if (count) *count = ct;
  return LPMX_OK;
}

}  // extern "C"
