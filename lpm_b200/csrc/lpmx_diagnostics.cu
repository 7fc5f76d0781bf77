// lpmx_diagnostics.cu -- the callers' per-step O(N) tail on the device: conserved totals of
// the Incompressible2D state and the weighted error norms the example drivers log every step.
//   Incompressible2D::total_vorticity / total_enstrophy / total_kinetic_energy   src/lpm_incompressible2d_impl.hpp:91-137
//   ErrNorms / ReduceErrorFtor                                                  src/lpm_error.hpp:81-131, src/lpm_error_impl.hpp:59-108
//   ComputeFTLE<Seed> (quadrilateral faces, sphere and plane), get_max_ftle      src/mesh/lpm_ftle.hpp:15-338
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


// block-wide maximum with the reference's comparison (m > x ? m : x), identity = lowest double
__device__ __forceinline__ void block_combine_max1(double* v) {
  __shared__ double shm[kRedThreads];
  shm[threadIdx.x] = v[0];
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) shm[threadIdx.x] = shm[threadIdx.x] > shm[threadIdx.x + s] ? shm[threadIdx.x] : shm[threadIdx.x + s];
    __syncthreads();
  }
  v[0] = shm[0];
}

// ---- FTLE (src/mesh/lpm_ftle.hpp), one thread per face, as coded ---------------------------------------------------
// north_pole_rotation_matrix (src/util/lpm_math.hpp:199-217)
__device__ __forceinline__ void north_pole_rotation(double* r, const double* x) {
  const double cosy = sqrt(x[1] * x[1] + x[2] * x[2]);
  const double siny = x[0];
  const bool on_x_axis = fabs(cosy) < 2.220446049250313e-16;
  const double cosx = on_x_axis ? 1.0 : x[2] / cosy;
  const double sinx = on_x_axis ? 0.0 : x[1] / cosy;
  r[0] = cosy, r[1] = -sinx * siny, r[2] = -cosx * siny;
  r[3] = 0.0, r[4] = cosx, r[5] = -sinx;
  r[6] = siny, r[7] = cosy * sinx, r[8] = cosx * cosy;
}

// set_flow_map_gradient, cauchy_green_tensor (the elementwise product F_ij F_ji, :70-79) and the larger root of
// two_by_two_real_eigenvalues (src/util/lpm_math.hpp:119-141); returns log(lambda_1) (:218, no division by 2t)
__device__ __forceinline__ double ftle_log_lambda(const double* e0p, const double* e1p, const double* xd, const double* yd,
                                                  double dx0, double dy0) {
  const double F0 = (e1p[0] * xd[0] + e1p[1] * xd[1]) / dx0;
  const double F1 = (e0p[0] * xd[0] + e0p[1] * xd[1]) / dx0;
  const double F2 = (e1p[0] * yd[0] + e1p[1] * yd[1]) / dy0;
  const double F3 = (e0p[0] * yd[0] + e0p[1] * yd[1]) / dy0;
  const double c0 = F0 * F0, c1 = F1 * F2, c2 = F2 * F1, c3 = F3 * F3;
  const double det = c0 * c3 - c1 * c2;
  const double half_trace = 0.5 * (c0 + c3);
  double sqrt_arg = half_trace * half_trace - det;
  if (fabs(sqrt_arg) < 2.220446049250313e-16) sqrt_arg = 0.0;
  return log(half_trace + sqrt(sqrt_arg));
}

struct IntQuadView {  // Kokkos::View<Index*[4]> in either layout
  const int* p;
  long si, sk;
  __device__ int operator()(long i, int k) const { return p[i * si + k * sk]; }
};

template <int NDIM>
__global__ void __launch_bounds__(kRedThreads) ftle_kernel(int n_faces, Vec3View vert_phys, Vec3View vert_ref,
                                                           Vec3View face_phys, Vec3View face_ref, IntQuadView face_verts,
                                                           const unsigned char* __restrict__ mask,
                                                           double* __restrict__ ftle, double* __restrict__ part) {
  double vmax[1] = {-1.7976931348623157e308};  // Kokkos::Max identity
  for (long f = blockIdx.x * (long)kRedThreads + threadIdx.x; f < n_faces; f += (long)gridDim.x * kRedThreads) {
    if (mask[f]) continue;  // divided faces are skipped and their ftle entry is left alone (:90)
    double vp[4][2], vr[4][2];
    if (NDIM == 3) {
      double fa[3], fx[3], rr[9], rp[9];
      for (int k = 0; k < 3; ++k) fa[k] = face_ref(f, k), fx[k] = face_phys(f, k);
      // SphereGeometry::normalize(fxi) writes through the subview: the face's physical coordinates are normalised
      // in the caller's array (:98)
      const double s = 1.0 / sqrt(fx[0] * fx[0] + fx[1] * fx[1] + fx[2] * fx[2]);
      for (int k = 0; k < 3; ++k) {
        fx[k] *= s;
        face_phys(f, k) = fx[k];
      }
      north_pole_rotation(rr, fa);
      north_pole_rotation(rp, fx);
      for (int i = 0; i < 4; ++i) {
        const int v = face_verts(f, i);
        double xr[3], xp[3];
        for (int k = 0; k < 3; ++k) xr[k] = vert_ref(v, k), xp[k] = vert_phys(v, k);
        for (int a = 0; a < 2; ++a) {  // apply_3by3, rows 0 and 1 (the third tangent coordinate is never read)
          double tr = 0.0, tp = 0.0;
          for (int k = 0; k < 3; ++k) {
            tr += rr[3 * a + k] * xr[k];
            tp += rp[3 * a + k] * xp[k];
          }
          vr[i][a] = tr, vp[i][a] = tp;
        }
      }
    } else {
      for (int i = 0; i < 4; ++i) {
        const int v = face_verts(f, i);
        for (int a = 0; a < 2; ++a) vr[i][a] = vert_ref(v, a), vp[i][a] = vert_phys(v, a);
      }
    }
    // "shift so that vertex 1 is the origin" (:148-153, :254-259) runs in place over i = 0..3: vertex 0 is shifted,
    // vertex 1 becomes 0, vertices 2 and 3 are then shifted by 0.  Kept as coded.
    for (int a = 0; a < 2; ++a) {
      vp[0][a] -= vp[1][a];
      vr[0][a] -= vr[1][a];
      vp[1][a] = 0.0;
      vr[1][a] = 0.0;
    }
    const double e1r[2] = {vr[2][0] - vr[1][0], vr[2][1] - vr[1][1]};
    const double e1p[2] = {vp[2][0] - vp[1][0], vp[2][1] - vp[1][1]};
    const double dx0 = sqrt(e1r[0] * e1r[0] + e1r[1] * e1r[1]);
    double xd[2] = {e1r[0] / dx0, e1r[1] / dx0};
    const double e0r[2] = {vr[0][0], vr[0][1]}, e0p[2] = {vp[0][0], vp[0][1]};
    double yd[2], dy0;
    if (NDIM == 3) {  // orthogonalise against x (:181-199)
      const double dxy = xd[0] * e0r[0] + xd[1] * e0r[1];
      yd[0] = e0r[0] - dxy * xd[0], yd[1] = e0r[1] - dxy * xd[1];
    } else {  // the plane branch takes the reverse of edge 0 as is (:283-295)
      yd[0] = e0r[0], yd[1] = e0r[1];
    }
    dy0 = sqrt(yd[0] * yd[0] + yd[1] * yd[1]);
    yd[0] /= dy0, yd[1] /= dy0;
    const double val = ftle_log_lambda(e0p, e1p, xd, yd, dx0, dy0);
    ftle[f] = val;
    vmax[0] = vmax[0] > val ? vmax[0] : val;  // get_max_ftle's comparison (a NaN entry makes the reference's maximum order dependent too)
  }
  block_combine_max1(vmax);
  if (threadIdx.x == 0) part[blockIdx.x] = vmax[0];
}

__global__ void __launch_bounds__(kRedThreads) ftle_max_kernel(const double* __restrict__ part, int n_blocks,
                                                               double* __restrict__ out) {
  double v[1] = {-1.7976931348623157e308};
  for (int b = threadIdx.x; b < n_blocks; b += kRedThreads) v[0] = v[0] > part[b] ? v[0] : part[b];
  block_combine_max1(v);
  if (threadIdx.x == 0) out[0] = v[0];
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

int lpmx_ftle(lpmx_handle_t h, int geom, int n_verts, const double* vert_phys, const double* vert_ref, int layout,
              long vert_ld, int n_faces, double* face_phys, const double* face_ref, long face_ld, const int* face_verts,
              int verts_layout, const unsigned char* face_mask, double* ftle, double* max_ftle) {
  if (!h) return LPMX_ERR_INVALID;
  if (geom != LPMX_GEOM_SPHERE && geom != LPMX_GEOM_PLANE) return set_error(h, LPMX_ERR_INVALID, "unknown geometry");
  if (n_verts < 0 || n_faces < 0) return set_error(h, LPMX_ERR_INVALID, "negative extent");
  if (n_faces > 0 && (!vert_phys || !vert_ref || !face_phys || !face_ref || !face_verts || !face_mask || !ftle))
    return set_error(h, LPMX_ERR_INVALID, "null array");
  if ((layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) ||
      (verts_layout != LPMX_LAYOUT_LEFT && verts_layout != LPMX_LAYOUT_RIGHT))
    return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (layout == LPMX_LAYOUT_LEFT && (vert_ld < n_verts || face_ld < n_faces))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const int nd = geom == LPMX_GEOM_SPHERE ? 3 : 2;
  auto vb = [&](long ld, int n) {
    return (layout == LPMX_LAYOUT_LEFT ? (size_t)((nd - 1) * ld + n) : (size_t)nd * n) * sizeof(double);
  };
  double mx = -1.7976931348623157e308;
  if (n_faces > 0) {
    const void *dvp, *dvr, *dfr, *dfv, *dm;
    void *dfp, *dout;
    LPMX_TRY(stage_in(h, "ftle_vp", vert_phys, vb(vert_ld, n_verts), &dvp));
    LPMX_TRY(stage_in(h, "ftle_vr", vert_ref, vb(vert_ld, n_verts), &dvr));
    LPMX_TRY(stage_in(h, "ftle_fr", face_ref, vb(face_ld, n_faces), &dfr));
    LPMX_TRY(stage_in(h, "ftle_fv", face_verts, sizeof(int) * 4 * (size_t)n_faces, &dfv));
    LPMX_TRY(stage_in(h, "ftle_m", face_mask, (size_t)n_faces, &dm));
    // in/out arguments: staged in, and copied back after the kernel when they live on the host
    const void* tmp;
    LPMX_TRY(stage_in(h, "ftle_fp", face_phys, vb(face_ld, n_faces), &tmp));
    dfp = const_cast<void*>(tmp);
    LPMX_TRY(stage_in(h, "ftle_out", ftle, sizeof(double) * (size_t)n_faces, &tmp));
    dout = const_cast<void*>(tmp);
    void* scratch = nullptr;
    LPMX_TRY(dev_buffer(h, "red_scratch", sizeof(double) * (6 * (size_t)kRedMaxBlocks + 8), &scratch));
    double* part = (double*)scratch;
    double* res = part + 6 * (size_t)kRedMaxBlocks;
    const int blocks = red_blocks(n_faces);
    auto view = [&](const void* p, long ld) {
      Vec3View v;
      v.p = (double*)const_cast<void*>(p);
      if (layout == LPMX_LAYOUT_LEFT) v.si = 1, v.sk = ld;
      else v.si = nd, v.sk = 1;
      return v;
    };
    IntQuadView fv;
    fv.p = (const int*)dfv;
    if (verts_layout == LPMX_LAYOUT_LEFT) fv.si = 1, fv.sk = n_faces;
    else fv.si = 4, fv.sk = 1;
    if (nd == 3)
      ftle_kernel<3><<<blocks, kRedThreads, 0, h->stream>>>(n_faces, view(dvp, vert_ld), view(dvr, vert_ld), view(dfp, face_ld),
                                                            view(dfr, face_ld), fv, (const unsigned char*)dm, (double*)dout, part);
    else
      ftle_kernel<2><<<blocks, kRedThreads, 0, h->stream>>>(n_faces, view(dvp, vert_ld), view(dvr, vert_ld), view(dfp, face_ld),
                                                            view(dfr, face_ld), fv, (const unsigned char*)dm, (double*)dout, part);
    ftle_max_kernel<<<1, kRedThreads, 0, h->stream>>>(part, blocks, res);
    h->launches += 2;
    LPMX_CUDA(h, cudaGetLastError());
    if (dfp != (void*)face_phys && nd == 3) LPMX_TRY(stage_out_end(h, face_phys, dfp, vb(face_ld, n_faces)));
    if (dout != (void*)ftle) LPMX_TRY(stage_out_end(h, ftle, dout, sizeof(double) * (size_t)n_faces));
    LPMX_CUDA(h, cudaMemcpyAsync(&mx, res, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if (max_ftle) *max_ftle = mx;
  return LPMX_OK;
}

}  // extern "C"
