// lpmx_gmls.cu -- the steps either side of the spherical SWE direct sums, on the device:
//   GatherMeshData / ScatterMeshData       src/mesh/lpm_gather_mesh_data_impl.hpp:14-66, src/mesh/lpm_scatter_mesh_data_impl.hpp
//   gmls::Neighborhoods                    src/lpm_compadre.cpp:75-112 (k-nearest + window search, there on a host kd-tree)
//   gmls::sphere_scalar_gmls + Evaluator   src/lpm_compadre.hpp:163-195, call sites src/lpm_swe_rk2_impl.hpp:56-77,134-154,233-252
// The reference copies the gathered particles to the host every RK stage, searches a kd-tree and runs Compadre; here the
// particles never leave HBM: a uniform-grid sort (CUB radix sort on the cell index), one thread per target for the
// search + weighted least squares (lpmx_gmls_core.h), and the scatter back to vertex / face fields.  All kernels are
// O(N) and HBM/latency bound; they are < 2 % of a SWERK2 step (profiles/README.md).
// Compadre itself is absent from the reference tree: parity for the Laplacian values is UNPINNED (see lpmx_gmls_core.h).
#include <cub/cub.cuh>

#include "lpmx_gmls_core.h"
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

// strided n x ncomp accessor (both Kokkos layouts)
struct FieldView {
  double* p;
  long si, sk;
  __device__ double& operator()(long i, int k) const { return p[i * si + k * sk]; }
};
static FieldView field_view(const void* p, int layout, long ld, int ncomp) {
  FieldView v;
  v.p = (double*)const_cast<void*>(p);
  if (layout == LPMX_LAYOUT_LEFT) v.si = 1, v.sk = ld;
  else v.si = ncomp, v.sk = 1;
  return v;
}

// GatherScalarFaceData / GatherVectorFaceData + the vertex copy: row i < nv <- vertex i; row nv + leaf_idx(f) <- leaf f
__global__ void gather_kernel(int nv, int nf, int ncomp, FieldView vert, FieldView face, const unsigned char* __restrict__ mask,
                              const int* __restrict__ leaf_idx, FieldView out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < nv) {
    for (int k = 0; k < ncomp; ++k) out(i, k) = vert(i, k);
  } else if (i < (long)nv + nf) {
    const long f = i - nv;
    if (!mask[f]) {
      const long o = nv + leaf_idx[f];
      for (int k = 0; k < ncomp; ++k) out(o, k) = face(f, k);
    }
  }
}
// ScatterMeshData: the inverse; divided faces are not written
__global__ void scatter_kernel(int nv, int nf, int ncomp, FieldView gathered, const unsigned char* __restrict__ mask,
                               const int* __restrict__ leaf_idx, FieldView vert, FieldView face) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < nv) {
    for (int k = 0; k < ncomp; ++k) vert(i, k) = gathered(i, k);
  } else if (i < (long)nv + nf) {
    const long f = i - nv;
    if (!mask[f]) {
      const long o = nv + leaf_idx[f];
      for (int k = 0; k < ncomp; ++k) face(f, k) = gathered(o, k);
    }
  }
}

__global__ void max_radius2_kernel(int n, FieldView x, unsigned long long* out) {
  double m = 0.0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double r2 = x(i, 0) * x(i, 0) + x(i, 1) * x(i, 1) + x(i, 2) * x(i, 2);
    m = r2 > m ? r2 : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, m, o);
    m = v > m ? v : m;
  }
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void cell_key_kernel(int n, FieldView x, gmls::Cloud c, unsigned int* __restrict__ key, int* __restrict__ idx) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int a = gmls::cell_coord(c, x(i, 0)), b = gmls::cell_coord(c, x(i, 1)), d = gmls::cell_coord(c, x(i, 2));
  key[i] = (a * (unsigned int)c.G + b) * (unsigned int)c.G + d;
  idx[i] = (int)i;
}

__global__ void permute_xyz_kernel(int n, const int* __restrict__ idx, FieldView x, double* __restrict__ xs) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[i];
  xs[i] = x(j, 0);
  xs[(long)n + i] = x(j, 1);
  xs[2L * n + i] = x(j, 2);
}
__global__ void permute_scalar_kernel(int n, const int* __restrict__ idx, const double* __restrict__ f, double* __restrict__ fs) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) fs[i] = f[idx[i]];
}

// cell_start[q] = first sorted position whose key is >= q (q = 0..ncell)
__global__ void cell_start_kernel(long ncell, int n, const unsigned int* __restrict__ key_sorted, int* __restrict__ cell_start) {
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (q > ncell) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long)key_sorted[mid] < q) lo = mid + 1;
    else hi = mid;
  }
  cell_start[q] = lo;
}

template <int OM, int KMAX>
__global__ void __launch_bounds__(128) gmls_laplacian_kernel(gmls::Cloud c, gmls::Params p, const int* __restrict__ idx,
                                                             double* __restrict__ lap, double* __restrict__ eps_out,
                                                             int* __restrict__ nn_out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const gmls::TargetResult r = gmls::laplacian_at_target<OM, KMAX>(c, p, (int)i);
  const int o = idx[i];
  lap[o] = r.lap;
  if (eps_out) eps_out[o] = r.eps;
  if (nn_out) nn_out[o] = r.n_neighbors;
}

struct InterpOut {
  double* p[gmls::kInterpFields];
};
template <int OM, int KMAX>
__global__ void __launch_bounds__(128) gmls_interpolate_kernel(gmls::Cloud c, gmls::Fields fl, gmls::Params p, int n_tgt, FieldView xt,
                                                               InterpOut out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_tgt) return;
  double v[gmls::kInterpFields];
  gmls::interpolate_at_point<OM, KMAX>(c, fl, p, xt(i, 0), xt(i, 1), xt(i, 2), v);
#pragma unroll
  for (int q = 0; q < gmls::kInterpFields; ++q)
    if (out.p[q]) out.p[q][i] = v[q];
}

template <int OM>
static void launch_gmls(cudaStream_t st, const gmls::Cloud& c, const gmls::Params& p, const int* idx, double* lap, double* eps_out,
                        int* nn_out) {
  const unsigned blocks = (unsigned)((c.n + 127) / 128);
  if (p.min_neighbors <= 16)
    gmls_laplacian_kernel<OM, 16><<<blocks, 128, 0, st>>>(c, p, idx, lap, eps_out, nn_out);
  else
    gmls_laplacian_kernel<OM, gmls::kMaxK><<<blocks, 128, 0, st>>>(c, p, idx, lap, eps_out, nn_out);
}

static int check_params(lpmx_handle_t h, const lpmx_gmls_params_t* q, gmls::Params* p) {
  if (!q) return set_error(h, LPMX_ERR_INVALID, "null gmls params");
  if (q->samples_order < 2 || q->samples_order > gmls::kMaxOrder || q->manifold_order < 1 || q->manifold_order > gmls::kMaxOrder)
    return set_error(h, LPMX_ERR_UNSUPPORTED, "gmls orders must be in [2, %d] (samples) and [1, %d] (manifold)", gmls::kMaxOrder,
                     gmls::kMaxOrder);
  if (q->samples_weight_pwr != q->manifold_weight_pwr)
    return set_error(h, LPMX_ERR_UNSUPPORTED, "samples_weight_pwr != manifold_weight_pwr is not supported (gmls::Params sets both to 2)");
  if (q->ambient_dim != 3 || q->topo_dim != 2) return set_error(h, LPMX_ERR_UNSUPPORTED, "only ambient_dim 3 / topo_dim 2 (sphere)");
  if (q->min_neighbors < 3 || q->min_neighbors > gmls::kMaxK) return set_error(h, LPMX_ERR_UNSUPPORTED, "min_neighbors must be in [3, %d]", gmls::kMaxK);
  if (!(q->eps_multiplier >= 1.0) || !(q->samples_weight_pwr > 0.0)) return set_error(h, LPMX_ERR_INVALID, "bad eps_multiplier / weight power");
  p->samples_order = q->samples_order, p->manifold_order = q->manifold_order, p->min_neighbors = q->min_neighbors;
  p->eps_multiplier = q->eps_multiplier, p->weight_pwr = q->samples_weight_pwr;
  return LPMX_OK;
}

// Sort the cloud by grid cell: fills c (sorted coordinates, cell table; c.f is left null) and returns the sorted -> original
// index map.  One 8-byte read back (the bounding radius sizes the grid on the host).
static int build_cloud(lpmx_handle_t h, const gmls::Params& p, int n, FieldView x, gmls::Cloud* cloud, const int** perm) {
  const int threads = 256, blocks = (n + threads - 1) / threads;
  void* d_r = nullptr;
  LPMX_TRY(dev_buffer(h, "gmls_radius", 8, &d_r));
  LPMX_CUDA(h, cudaMemsetAsync(d_r, 0, 8, h->stream));
  max_radius2_kernel<<<blocks < 592 ? blocks : 592, threads, 0, h->stream>>>(n, x, (unsigned long long*)d_r);
  ++h->launches;
  double r2 = 0.0;
  LPMX_CUDA(h, cudaMemcpyAsync(&r2, d_r, 8, cudaMemcpyDeviceToHost, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (!(r2 > 0.0) || !std::isfinite(r2)) return set_error(h, LPMX_ERR_INVALID, "gmls: coordinates are zero or not finite");
  const gmls::GridDims gd = gmls::grid_dims(n, p.min_neighbors, p.eps_multiplier, sqrt(r2));
  gmls::Cloud c;
  c.n = n, c.G = gd.G, c.box = gd.box, c.cell = gd.cell, c.inv_cell = 1.0 / gd.cell;
  const long ncell = (long)c.G * c.G * c.G;
  void *d_key = nullptr, *d_key2 = nullptr, *d_idx = nullptr, *d_idx2 = nullptr, *d_xs = nullptr, *d_cs = nullptr, *d_tmp = nullptr;
  LPMX_TRY(dev_buffer(h, "gmls_key", 4 * (size_t)n, &d_key));
  LPMX_TRY(dev_buffer(h, "gmls_key2", 4 * (size_t)n, &d_key2));
  LPMX_TRY(dev_buffer(h, "gmls_idx", 4 * (size_t)n, &d_idx));
  LPMX_TRY(dev_buffer(h, "gmls_idx2", 4 * (size_t)n, &d_idx2));
  LPMX_TRY(dev_buffer(h, "gmls_xs", 8 * 3 * (size_t)n, &d_xs));
  LPMX_TRY(dev_buffer(h, "gmls_cell_start", 4 * (size_t)(ncell + 1), &d_cs));
  cell_key_kernel<<<blocks, threads, 0, h->stream>>>(n, x, c, (unsigned int*)d_key, (int*)d_idx);
  int bits = 1;
  while ((1L << bits) < ncell) ++bits;
  size_t tmp_bytes = 0;
  LPMX_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned int*)d_key, (unsigned int*)d_key2,
                                               (const int*)d_idx, (int*)d_idx2, n, 0, bits, h->stream));
  LPMX_TRY(dev_buffer(h, "gmls_sort_tmp", tmp_bytes + 16, &d_tmp));
  LPMX_CUDA(h, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, (const unsigned int*)d_key, (unsigned int*)d_key2,
                                               (const int*)d_idx, (int*)d_idx2, n, 0, bits, h->stream));
  permute_xyz_kernel<<<blocks, threads, 0, h->stream>>>(n, (const int*)d_idx2, x, (double*)d_xs);
  cell_start_kernel<<<(unsigned)((ncell + 1 + threads - 1) / threads), threads, 0, h->stream>>>(ncell, n, (const unsigned int*)d_key2,
                                                                                                (int*)d_cs);
  h->launches += 4;  // keys, sort (counted once), permute, cell table
  LPMX_CUDA(h, cudaGetLastError());
  c.x = (const double*)d_xs, c.f = nullptr, c.cell_start = (const int*)d_cs;
  *cloud = c;
  *perm = (const int*)d_idx2;
  return LPMX_OK;
}

// x: n x 3 device view; f, lap: device arrays of n; eps_out / nn_out optional device arrays
static int gmls_laplacian_device(lpmx_handle_t h, const gmls::Params& p, int n, FieldView x, const double* f, double* lap,
                                 double* eps_out, int* nn_out) {
  if (n <= 0) return LPMX_OK;
  gmls::Cloud c;
  const int* perm;
  LPMX_TRY(build_cloud(h, p, n, x, &c, &perm));
  void* d_fs = nullptr;
  LPMX_TRY(dev_buffer(h, "gmls_fs", 8 * (size_t)n, &d_fs));
  permute_scalar_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, perm, f, (double*)d_fs);
  c.f = (const double*)d_fs;
  const int om = p.samples_order > p.manifold_order ? p.samples_order : p.manifold_order;
  if (om <= 2) launch_gmls<2>(h->stream, c, p, perm, lap, eps_out, nn_out);
  else if (om == 3) launch_gmls<3>(h->stream, c, p, perm, lap, eps_out, nn_out);
  else launch_gmls<4>(h->stream, c, p, perm, lap, eps_out, nn_out);
  h->launches += 2;
  LPMX_CUDA(h, cudaGetLastError());
  return LPMX_OK;
}

// scalar point evaluation of n_fields source fields (device arrays of n_src) at n_tgt target points -> out[f] (device, n_tgt)
static int gmls_interpolate_device(lpmx_handle_t h, const gmls::Params& p, int n_src, FieldView xs, int n_fields,
                                   const double* const* fields, int n_tgt, FieldView xt, double* const* out) {
  if (n_tgt <= 0 || n_fields <= 0) return LPMX_OK;
  gmls::Cloud c;
  const int* perm;
  LPMX_TRY(build_cloud(h, p, n_src, xs, &c, &perm));
  void* d_fs = nullptr;
  LPMX_TRY(dev_buffer(h, "gmls_fs4", 8 * (size_t)gmls::kInterpFields * (size_t)n_src, &d_fs));
  const int om = p.samples_order < 2 ? 2 : p.samples_order;
  for (int f0 = 0; f0 < n_fields; f0 += gmls::kInterpFields) {
    gmls::Fields fl;
    InterpOut o;
    for (int q = 0; q < gmls::kInterpFields; ++q) {
      const int f = f0 + q < n_fields ? f0 + q : f0;  // unused slots repeat the first field of the batch
      double* dst = (double*)d_fs + (size_t)q * n_src;
      if (f0 + q < n_fields || q == 0) {
        permute_scalar_kernel<<<(n_src + 255) / 256, 256, 0, h->stream>>>(n_src, perm, fields[f], dst);
        ++h->launches;
        fl.f[q] = dst;
      } else {
        fl.f[q] = (const double*)d_fs;
      }
      o.p[q] = f0 + q < n_fields ? out[f0 + q] : nullptr;
    }
    const unsigned blocks = (unsigned)((n_tgt + 127) / 128);
    const bool small = p.min_neighbors <= 16;
    if (om == 2) {
      if (small) gmls_interpolate_kernel<2, 16><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
      else gmls_interpolate_kernel<2, gmls::kMaxK><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
    } else if (om == 3) {
      if (small) gmls_interpolate_kernel<3, 16><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
      else gmls_interpolate_kernel<3, gmls::kMaxK><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
    } else {
      if (small) gmls_interpolate_kernel<4, 16><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
      else gmls_interpolate_kernel<4, gmls::kMaxK><<<blocks, 128, 0, h->stream>>>(c, fl, p, n_tgt, xt, o);
    }
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  return LPMX_OK;
}

static int leaf_index(lpmx_handle_t h, const unsigned char* d_mask, int nf, int** d_leaf, int* n_leaf) {
  void* p = nullptr;
  LPMX_TRY(dev_buffer(h, "gmls_leaf_idx", sizeof(int) * (size_t)(nf + 1), &p));
  *d_leaf = (int*)p;
  *n_leaf = 0;
  if (nf > 0) LPMX_TRY(scan_leaves(h, d_mask, nf, *d_leaf, n_leaf));
  return LPMX_OK;
}

}  // namespace lpmx

extern "C" {

int lpmx_gmls_params_init(lpmx_gmls_params_t* p, int order) {
  if (!p || order < 1) return LPMX_ERR_INVALID;
  // gmls::Params(order, dim = 3) (src/lpm_compadre.hpp:49-60)
  p->eps_multiplier = 2.0;
  p->samples_order = order;
  p->manifold_order = order;
  p->samples_weight_pwr = 2.0;
  p->manifold_weight_pwr = 2.0;
  p->ambient_dim = 3;
  p->topo_dim = 2;
  p->min_neighbors = (order + 1) * (order + 2) / 2;  // Compadre::GMLS::getNP(order, topo_dim)
  return LPMX_OK;
}

static size_t field_bytes(int layout, long ld, int n, int ncomp) {
  return (layout == LPMX_LAYOUT_LEFT ? (size_t)((ncomp - 1) * ld + n) : (size_t)ncomp * n) * sizeof(double);
}

int lpmx_gather_mesh_data(lpmx_handle_t h, int n_comp, int layout, int n_verts, const double* vert_data, long vert_ld, int n_faces,
                          const double* face_data, long face_ld, const unsigned char* face_mask, double* gathered, long gathered_ld,
                          int* n_gathered) {
  if (!h) return LPMX_ERR_INVALID;
  if (n_comp < 1 || n_comp > 3 || n_verts < 0 || n_faces < 0) return set_error(h, LPMX_ERR_INVALID, "bad extent");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((n_verts > 0 && !vert_data) || (n_faces > 0 && (!face_data || !face_mask))) return set_error(h, LPMX_ERR_INVALID, "null array");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const void* dm;
  LPMX_TRY(stage_in(h, "gs_mask", face_mask, (size_t)n_faces, &dm));
  int *d_leaf, n_leaf;
  LPMX_TRY(leaf_index(h, (const unsigned char*)dm, n_faces, &d_leaf, &n_leaf));
  const int n = n_verts + n_leaf;
  if (n_gathered) *n_gathered = n;
  if (!gathered) return LPMX_OK;  // size query
  if (layout == LPMX_LAYOUT_LEFT && (vert_ld < n_verts || face_ld < n_faces || gathered_ld < n))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  const void *dv, *df;
  void* dout;
  LPMX_TRY(stage_in(h, "gs_vert", vert_data, field_bytes(layout, vert_ld, n_verts, n_comp), &dv));
  LPMX_TRY(stage_in(h, "gs_face", face_data, field_bytes(layout, face_ld, n_faces, n_comp), &df));
  LPMX_TRY(stage_out_begin(h, "gs_out", gathered, field_bytes(layout, gathered_ld, n, n_comp), &dout));
  const long nt = (long)n_verts + n_faces;
  if (nt > 0) {
    gather_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, h->stream>>>(
        n_verts, n_faces, n_comp, field_view(dv, layout, vert_ld, n_comp), field_view(df, layout, face_ld, n_comp),
        (const unsigned char*)dm, d_leaf, field_view(dout, layout, gathered_ld, n_comp));
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(stage_out_end(h, gathered, dout, field_bytes(layout, gathered_ld, n, n_comp)));
  if (dout != (void*)gathered) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

int lpmx_scatter_mesh_data(lpmx_handle_t h, int n_comp, int layout, const double* gathered, long gathered_ld, int n_verts,
                           double* vert_data, long vert_ld, int n_faces, double* face_data, long face_ld,
                           const unsigned char* face_mask) {
  if (!h) return LPMX_ERR_INVALID;
  if (n_comp < 1 || n_comp > 3 || n_verts < 0 || n_faces < 0) return set_error(h, LPMX_ERR_INVALID, "bad extent");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (!gathered || (n_verts > 0 && !vert_data) || (n_faces > 0 && (!face_data || !face_mask)))
    return set_error(h, LPMX_ERR_INVALID, "null array");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const void* dm;
  LPMX_TRY(stage_in(h, "gs_mask", face_mask, (size_t)n_faces, &dm));
  int *d_leaf, n_leaf;
  LPMX_TRY(leaf_index(h, (const unsigned char*)dm, n_faces, &d_leaf, &n_leaf));
  const int n = n_verts + n_leaf;
  if (layout == LPMX_LAYOUT_LEFT && (vert_ld < n_verts || face_ld < n_faces || gathered_ld < n))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  const void *dg, *tmp;
  LPMX_TRY(stage_in(h, "gs_out", gathered, field_bytes(layout, gathered_ld, n, n_comp), &dg));
  // face_data is in/out (divided faces keep their values): stage it in, copy all of it back
  LPMX_TRY(stage_in(h, "gs_face", face_data, field_bytes(layout, face_ld, n_faces, n_comp), &tmp));
  void* df = const_cast<void*>(tmp);
  void* dv;
  LPMX_TRY(stage_out_begin(h, "gs_vert", vert_data, field_bytes(layout, vert_ld, n_verts, n_comp), &dv));
  const long nt = (long)n_verts + n_faces;
  if (nt > 0) {
    scatter_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, h->stream>>>(
        n_verts, n_faces, n_comp, field_view(dg, layout, gathered_ld, n_comp), (const unsigned char*)dm, d_leaf,
        field_view(dv, layout, vert_ld, n_comp), field_view(df, layout, face_ld, n_comp));
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(stage_out_end(h, vert_data, dv, field_bytes(layout, vert_ld, n_verts, n_comp)));
  LPMX_TRY(stage_out_end(h, face_data, df, field_bytes(layout, face_ld, n_faces, n_comp)));
  if (dv != (void*)vert_data || df != (void*)face_data) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

int lpmx_gmls_sphere_laplacian(lpmx_handle_t h, const lpmx_gmls_params_t* params, int n, const double* xyz, int layout, long ld,
                               const double* f, double* laplacian, double* window_radius, int* n_neighbors) {
  if (!h) return LPMX_ERR_INVALID;
  gmls::Params p;
  LPMX_TRY(check_params(h, params, &p));
  if (n < 0 || (n > 0 && (!xyz || !f || !laplacian))) return set_error(h, LPMX_ERR_INVALID, "null array");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (layout == LPMX_LAYOUT_LEFT && ld < n) return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  if (n == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const void *dx, *df;
  void *dl, *de = nullptr, *dn = nullptr;
  LPMX_TRY(stage_in(h, "gm_x", xyz, field_bytes(layout, ld, n, 3), &dx));
  LPMX_TRY(stage_in(h, "gm_f", f, sizeof(double) * (size_t)n, &df));
  LPMX_TRY(stage_out_begin(h, "gm_lap", laplacian, sizeof(double) * (size_t)n, &dl));
  if (window_radius) LPMX_TRY(stage_out_begin(h, "gm_eps", window_radius, sizeof(double) * (size_t)n, &de));
  if (n_neighbors) LPMX_TRY(stage_out_begin(h, "gm_nn", n_neighbors, sizeof(int) * (size_t)n, &dn));
  LPMX_TRY(gmls_laplacian_device(h, p, n, field_view(dx, layout, ld, 3), (const double*)df, (double*)dl, (double*)de, (int*)dn));
  LPMX_TRY(stage_out_end(h, laplacian, dl, sizeof(double) * (size_t)n));
  LPMX_TRY(stage_out_end(h, window_radius, de, sizeof(double) * (size_t)n));
  LPMX_TRY(stage_out_end(h, n_neighbors, dn, sizeof(int) * (size_t)n));
  if (dl != (void*)laplacian || (window_radius && de != (void*)window_radius) || (n_neighbors && dn != (void*)n_neighbors))
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

int lpmx_gmls_sphere_interpolate(lpmx_handle_t h, const lpmx_gmls_params_t* params, int n_src, const double* src_xyz, int src_layout,
                                 long src_ld, int n_fields, const double* const* src_fields, int n_tgt, const double* tgt_xyz,
                                 int tgt_layout, long tgt_ld, double* const* tgt_fields) {
  if (!h) return LPMX_ERR_INVALID;
  lpmx_gmls_params_t q;
  if (!params) return set_error(h, LPMX_ERR_INVALID, "null gmls params");
  q = *params;
  if (q.manifold_order < 1) q.manifold_order = 1;
  gmls::Params p;
  if (q.samples_order == 1) {  // a linear fit is allowed for point evaluation
    q.samples_order = 2;
    LPMX_TRY(check_params(h, &q, &p));
    p.samples_order = 1;
  } else {
    LPMX_TRY(check_params(h, &q, &p));
  }
  if (n_src < 0 || n_tgt < 0 || n_fields < 0 || n_fields > 64) return set_error(h, LPMX_ERR_INVALID, "bad extent");
  if (n_tgt == 0 || n_fields == 0) return LPMX_OK;
  if (n_src == 0) return set_error(h, LPMX_ERR_INVALID, "no source points");
  if (!src_xyz || !tgt_xyz || !src_fields || !tgt_fields) return set_error(h, LPMX_ERR_INVALID, "null array");
  for (int f = 0; f < n_fields; ++f)
    if (!src_fields[f] || !tgt_fields[f]) return set_error(h, LPMX_ERR_INVALID, "null field %d", f);
  if ((src_layout != LPMX_LAYOUT_LEFT && src_layout != LPMX_LAYOUT_RIGHT) || (tgt_layout != LPMX_LAYOUT_LEFT && tgt_layout != LPMX_LAYOUT_RIGHT))
    return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((src_layout == LPMX_LAYOUT_LEFT && src_ld < n_src) || (tgt_layout == LPMX_LAYOUT_LEFT && tgt_ld < n_tgt))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  const void *dxs, *dxt;
  LPMX_TRY(stage_in(h, "gi_xs", src_xyz, field_bytes(src_layout, src_ld, n_src, 3), &dxs));
  LPMX_TRY(stage_in(h, "gi_xt", tgt_xyz, field_bytes(tgt_layout, tgt_ld, n_tgt, 3), &dxt));
  std::vector<const double*> din(n_fields);
  std::vector<double*> dout(n_fields);
  bool any_host = false;
  for (int f = 0; f < n_fields; ++f) {
    const std::string ni = "gi_in" + std::to_string(f), no = "gi_out" + std::to_string(f);
    const void* d;
    LPMX_TRY(stage_in(h, ni.c_str(), src_fields[f], sizeof(double) * (size_t)n_src, &d));
    din[f] = (const double*)d;
    void* o;
    LPMX_TRY(stage_out_begin(h, no.c_str(), tgt_fields[f], sizeof(double) * (size_t)n_tgt, &o));
    dout[f] = (double*)o;
    any_host = any_host || (o != (void*)tgt_fields[f]);
  }
  LPMX_TRY(gmls_interpolate_device(h, p, n_src, field_view(dxs, src_layout, src_ld, 3), n_fields, din.data(), n_tgt,
                                   field_view(dxt, tgt_layout, tgt_ld, 3), dout.data()));
  for (int f = 0; f < n_fields; ++f) LPMX_TRY(stage_out_end(h, tgt_fields[f], dout[f], sizeof(double) * (size_t)n_tgt));
  if (any_host) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

// lpmx_swe_laplacian_fn: gather (vertices + leaf faces) -> GMLS Laplacian of the surface height -> scatter
int lpmx_gmls_swe_laplacian(void* user, int stage, void* cuda_stream, int n_passive, const double* passive_xyz,
                            const double* passive_surf, double* passive_laps, int n_active, const double* active_xyz,
                            const double* active_surf, const unsigned char* active_mask, double* active_laps, long xyz_ld) {
  (void)stage;
  lpmx_gmls_provider_t* pr = (lpmx_gmls_provider_t*)user;
  if (!pr || !pr->handle) return LPMX_ERR_INVALID;
  lpmx_handle_t h = pr->handle;
  if ((cudaStream_t)cuda_stream != h->stream) return set_error(h, LPMX_ERR_INVALID, "gmls provider: called on a foreign stream");
  gmls::Params p;
  LPMX_TRY(check_params(h, &pr->params, &p));
  int *d_leaf, n_leaf;
  LPMX_TRY(leaf_index(h, active_mask, n_active, &d_leaf, &n_leaf));
  const int n = n_passive + n_leaf;
  if (n == 0) return LPMX_OK;
  void *gx, *gf, *gl;
  LPMX_TRY(dev_buffer(h, "gmls_gx", 8 * 3 * (size_t)n, &gx));
  LPMX_TRY(dev_buffer(h, "gmls_gf", 8 * (size_t)n, &gf));
  LPMX_TRY(dev_buffer(h, "gmls_gl", 8 * (size_t)n, &gl));
  const long nt = (long)n_passive + n_active;
  const unsigned blocks = (unsigned)((nt + 255) / 256);
  gather_kernel<<<blocks, 256, 0, h->stream>>>(n_passive, n_active, 3, field_view(passive_xyz, LPMX_LAYOUT_LEFT, xyz_ld, 3),
                                               field_view(active_xyz, LPMX_LAYOUT_LEFT, xyz_ld, 3), active_mask, d_leaf,
                                               field_view(gx, LPMX_LAYOUT_LEFT, n, 3));
  gather_kernel<<<blocks, 256, 0, h->stream>>>(n_passive, n_active, 1, field_view(passive_surf, LPMX_LAYOUT_RIGHT, 0, 1),
                                               field_view(active_surf, LPMX_LAYOUT_RIGHT, 0, 1), active_mask, d_leaf,
                                               field_view(gf, LPMX_LAYOUT_RIGHT, 0, 1));
  h->launches += 2;
  LPMX_TRY(gmls_laplacian_device(h, p, n, field_view(gx, LPMX_LAYOUT_LEFT, n, 3), (const double*)gf, (double*)gl, nullptr, nullptr));
  scatter_kernel<<<blocks, 256, 0, h->stream>>>(n_passive, n_active, 1, field_view(gl, LPMX_LAYOUT_RIGHT, 0, 1), active_mask, d_leaf,
                                                field_view(passive_laps, LPMX_LAYOUT_RIGHT, 0, 1),
                                                field_view(active_laps, LPMX_LAYOUT_RIGHT, 0, 1));
  ++h->launches;
  LPMX_CUDA(h, cudaGetLastError());
  return LPMX_OK;
}

}  // extern "C"
