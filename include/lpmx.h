/* lpmx.h -- C ABI of the B200-native direct-sum engine for LPM's spherical particle solvers.
 *
 * This is the drop-in boundary.  pbosler/lpm has no FFI layer: its de-facto
 * operator interface is the Kokkos functor constructors and the time-stepper entry points, all of
 * which take Kokkos Views by value.  Every function below names the reference interface it stands
 * in for (file:line under /root/reference/src).  A maintainer's binding passes `view.data()`,
 * the extents and the layout of the View (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch types; every function returns an int
 *     (LPMX_OK == 0, negative = error) and never throws.  lpmx_last_error_string() gives detail.
 *   - Real = double, Index = int, mask = unsigned char (Kokkos View<bool*>), as in
 *     LpmConfig.h.in:31-32 and lpm_kokkos_defs.hpp:31-39.
 *   - every array argument may be a DEVICE pointer (Kokkos CUDA build: used in place, no copy)
 *     or a HOST pointer (Kokkos OpenMP/Serial build: staged through pinned memory; the copies are
 *     part of the call).  The library detects which with cudaPointerGetAttributes.
 *   - `layout` describes Real*[3] Views: LPMX_LAYOUT_RIGHT is x[i*3+k] (Kokkos LayoutRight, the
 *     host default), LPMX_LAYOUT_LEFT is x[k*ld+i] (Kokkos LayoutLeft, the CUDA default,
 *     lpm_geometry.hpp:261-263).  `ld` is the leading dimension (>= n) for LAYOUT_LEFT and is
 *     ignored for LAYOUT_RIGHT.
 *   - calls are asynchronous on the handle's stream when all pointers are device pointers;
 *     lpmx_sync() waits.  Calls with host pointers return after results are in the host buffers.
 *   - there is NO CPU fallback: without a CUDA device lpmx_create fails with LPMX_ERR_NO_DEVICE.
 *     Only the mesh generator (host code in the reference too) works without a GPU.
 */
#ifndef LPMX_H
#define LPMX_H

#ifdef __cplusplus
extern "C" {
#endif

#define LPMX_VERSION_MAJOR 0
#define LPMX_VERSION_MINOR 1

/* error codes */
#define LPMX_OK 0
#define LPMX_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unknown enum) */
#define LPMX_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed */
#define LPMX_ERR_NOMEM (-3)       /* host or device allocation failed */
#define LPMX_ERR_NO_DEVICE (-4)   /* no usable CUDA device: the engine has no CPU fallback */
#define LPMX_ERR_COMM (-5)        /* NCCL / multi-GPU exchange failed */
#define LPMX_ERR_UNSUPPORTED (-6) /* recognised but out-of-scope request */
#define LPMX_ERR_STATE (-7)       /* call sequence error (e.g. advance before set_state) */

/* layouts of Real*[3] views */
#define LPMX_LAYOUT_RIGHT 0 /* x[i*3+k]  */
#define LPMX_LAYOUT_LEFT 1  /* x[k*ld+i] */

/* mesh seeds (reference: src/mesh/lpm_mesh_seed.hpp:111-147) */
#define LPMX_SEED_ICOS_TRI_SPHERE 0
#define LPMX_SEED_CUBED_SPHERE 1
#define LPMX_SEED_QUAD_RECT 2 /* QuadRectSeed (:46-62): planar quadrilaterals on [-r, r]^2, free boundary; Real*[2] coordinates */
#define LPMX_SEED_TRI_HEX 3   /* TriHexSeed (:64-82): planar triangles on a hexagon of circumradius r, free boundary          */

typedef struct lpmx_handle_s* lpmx_handle_t;
typedef struct lpmx_mesh_s* lpmx_mesh_t;
typedef struct lpmx_bve_solver_s* lpmx_bve_solver_t;
typedef struct lpmx_ic2d_solver_s* lpmx_ic2d_solver_t;
typedef struct lpmx_swe_solver_s* lpmx_swe_solver_t;
typedef struct lpmx_plane_solver_s* lpmx_plane_solver_t;

const char* lpmx_version_string(void);
const char* lpmx_error_name(int code);

/* ------------------------------------------------------------------------------------------
 * Mesh generator (HOST; deterministic; no GPU needed).
 * Replaces PolyMesh2d<Seed>::tree_init (src/mesh/lpm_polymesh2d_impl.hpp:25-42) with
 * MeshSeed<Seed> (src/mesh/lpm_mesh_seed.cpp:10-18,20-206), FaceDivider<..>::divide
 * (src/mesh/lpm_faces_impl.hpp:284-431 tri, :433-574 quad) and Edges::divide
 * (src/mesh/lpm_edges.cpp:58-96) for the two spherical seeds and the two planar seeds (the dividers are generic in the
 * geometry: SphereGeometry / PlaneGeometry midpoint, barycenter, polygon_area, src/lpm_geometry.hpp:69-160,516-642).
 * Coordinate rows have ndim entries: 3 on the sphere, 2 in the plane.  Boundary edges of the planar seeds have right = -1.
 * ------------------------------------------------------------------------------------------ */

/* MeshSeed<Seed>::set_max_allocations (src/mesh/lpm_mesh_seed.cpp:266-279) */
int lpmx_mesh_max_allocations(int seed, int depth, int* n_verts, int* n_edges, int* n_faces);

/* Build the tree mesh of `depth` uniform refinements; seed coordinates are multiplied by `radius` (sphere radius; half
 * width of the square; circumradius of the hexagon). */
int lpmx_mesh_create(int seed, int depth, double radius, lpmx_mesh_t* mesh);
int lpmx_mesh_destroy(lpmx_mesh_t mesh);

/* counts after refinement: vertices.nh(), edges.nh(), faces.nh(), faces.n_leaves_host(),
 * edges.n_leaves_host(), vertices per face (3 or 4) */
int lpmx_mesh_sizes(lpmx_mesh_t mesh, int* n_verts, int* n_edges, int* n_faces, int* n_face_leaves,
                    int* n_edge_leaves, int* n_face_verts);

/* array ids for lpmx_mesh_array: element type and extent in the comment */
#define LPMX_MESH_VERT_XYZ 0       /* double[n_verts][ndim]  vertices.phys_crds (LayoutRight); ndim = 3 sphere, 2 plane (also ids 1, 9, 10) */
#define LPMX_MESH_VERT_LAG_XYZ 1   /* double[n_verts][3]  vertices.lag_crds                      */
#define LPMX_MESH_VERT_CRD_INDS 2  /* int[n_verts]        vertices.crd_inds                      */
#define LPMX_MESH_EDGE_ORIGS 3     /* int[n_edges]        edges.origs                            */
#define LPMX_MESH_EDGE_DESTS 4     /* int[n_edges]        edges.dests                            */
#define LPMX_MESH_EDGE_LEFTS 5     /* int[n_edges]        edges.lefts                            */
#define LPMX_MESH_EDGE_RIGHTS 6    /* int[n_edges]        edges.rights                           */
#define LPMX_MESH_EDGE_PARENTS 7   /* int[n_edges]        edges.parent                           */
#define LPMX_MESH_EDGE_KIDS 8      /* int[n_edges][2]     edges.kids                             */
#define LPMX_MESH_FACE_XYZ 9       /* double[n_faces][3]  faces.phys_crds                        */
#define LPMX_MESH_FACE_LAG_XYZ 10  /* double[n_faces][3]  faces.lag_crds                         */
#define LPMX_MESH_FACE_AREA 11     /* double[n_faces]     faces.area  (0 for divided faces)      */
#define LPMX_MESH_FACE_MASK 12     /* uchar[n_faces]      faces.mask  (1 for divided faces)      */
#define LPMX_MESH_FACE_VERTS 13    /* int[n_faces][nfv]   faces.verts                            */
#define LPMX_MESH_FACE_EDGES 14    /* int[n_faces][nfv]   faces.edges                            */
#define LPMX_MESH_FACE_CRD_INDS 15 /* int[n_faces]        faces.crd_inds                         */
#define LPMX_MESH_FACE_PARENT 16   /* int[n_faces]        faces.parent                           */
#define LPMX_MESH_FACE_KIDS 17     /* int[n_faces][4]     faces.kids                             */
#define LPMX_MESH_FACE_LEVEL 18    /* int[n_faces]        faces.level (root level defined as 1)  */
#define LPMX_MESH_FACE_LEAF_IDX 19 /* int[n_faces]        faces.leaf_idx (exclusive scan)        */

/* Returns a pointer (owned by the mesh, valid until lpmx_mesh_destroy) to the array and its
 * element count; *is_real = 1 for double arrays, 0 for int arrays, 2 for unsigned char. */
int lpmx_mesh_array(lpmx_mesh_t mesh, int array_id, const void** data, long* count, int* is_real);

/* Overwrite one of the four coordinate arrays (LPMX_MESH_VERT_XYZ, _VERT_LAG_XYZ, _FACE_XYZ, _FACE_LAG_XYZ) with the
 * caller's current values (`count` doubles, LayoutRight) -- the particles move between mesh construction and an adaptive
 * refinement, and the reference divides faces from the coordinates as they are then. */
int lpmx_mesh_update_array(lpmx_mesh_t mesh, int array_id, const double* data, long count);

/* outcomes of lpmx_mesh_divide_flagged_faces */
#define LPMX_AMR_DIVIDED_ALL 0   /* every flagged face was divided                                                   */
#define LPMX_AMR_NO_SPACE 1      /* nothing divided: flag count > (max_faces - n_faces) / 4 (the reference warns and returns) */
#define LPMX_AMR_LIMIT_REACHED 2 /* flagged faces with level > max_level were left alone                              */

/* PolyMesh2d<Seed>::divide_flagged_faces (src/mesh/lpm_polymesh2d_impl.hpp:124-173): divide, in index order, every face
 * i < n_faces_host() with flags[i] != 0 whose level is <= max_level (the reference passes init_depth + amr_limit; root
 * faces have level 1 here, see LPMX_MESH_FACE_LEVEL), then rescan the leaves (Faces::scan_leaves).  max_faces is
 * PolyMeshParameters::nmaxfaces (src/mesh/lpm_polymesh2d.hpp:68-71, allocations for depth + amr_buffer).  flags has
 * n_flags >= n_faces entries (only the first n_faces are read).  A flagged face that is already divided is an error
 * (LPMX_ERR_INVALID; the reference asserts).  Pointers obtained from lpmx_mesh_array are invalidated. */
int lpmx_mesh_divide_flagged_faces(lpmx_mesh_t mesh, const unsigned char* flags, int n_flags, int max_faces,
                                   int max_level, int* n_divided, int* outcome);

/* Mesh queries on the (possibly adaptively refined) tree, host code like the mesh itself; all restated as coded.
 * PolyMesh2d::get_leaf_edges_from_parent / ccw_edges_around_face / ccw_adjacent_faces (src/mesh/lpm_polymesh2d.hpp:277-308,
 * :316-340, :348-366): *n receives the length of the list, the first min(*n, cap) entries are written.  Neighbours across a free
 * boundary of a planar mesh are -1.  The reference's own known answers (tests/lpm_polymesh2d_function_tests.cpp:96-104,213-238)
 * are asserted in tests/test_mesh_queries.py. */
int lpmx_mesh_leaf_edges_from_parent(lpmx_mesh_t mesh, int parent_edge, int* list, int cap, int* n);
int lpmx_mesh_ccw_edges_around_face(lpmx_mesh_t mesh, int face, int* list, int cap, int* n);
int lpmx_mesh_ccw_adjacent_faces(lpmx_mesh_t mesh, int face, int* list, int cap, int* n);
/* NeighborsFlag (src/mesh/lpm_refinement_flags.hpp:30-52): flags[i] |= (some neighbour of face i is more than one level
 * finer), for i in [start, end) -- what keeps an adaptive mesh 2:1 balanced.  *n_flagged counts the newly set flags. */
int lpmx_mesh_neighbors_flag(lpmx_mesh_t mesh, unsigned char* flags, int start, int end, int* n_flagged);
/* Point location at the mesh's current coordinates (push them with lpmx_mesh_update_array first when the particles have moved).
 * pts: n_pts rows of ndim doubles; out: face indices.  Modes: locate_face_containing_pt (:541-552; -1 when a planar point lies
 * outside the mesh, pt_is_outside_mesh :484-535), locate_pt_walk_search from the leaf start[i] (:377-413), locate_pt_tree_search
 * from the face start[i] (:446-473), nearest_root_face (:421-434). */
#define LPMX_LOCATE_CONTAINING 0
#define LPMX_LOCATE_WALK 1
#define LPMX_LOCATE_TREE 2
#define LPMX_LOCATE_NEAREST_ROOT 3
int lpmx_mesh_locate(lpmx_mesh_t mesh, int mode, const double* pts, int n_pts, const int* start, int* out);

/* ------------------------------------------------------------------------------------------
 * Engine handle: one per process and GPU.
 * ------------------------------------------------------------------------------------------ */
int lpmx_create(lpmx_handle_t* h, int device_id);
int lpmx_destroy(lpmx_handle_t h);
int lpmx_sync(lpmx_handle_t h);
const char* lpmx_last_error_string(lpmx_handle_t h);
/* cudaStream_t the handle launches on (as void*), for callers that time with CUDA events */
int lpmx_stream(lpmx_handle_t h, void** cuda_stream);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches claim) */
int lpmx_launch_count(lpmx_handle_t h, long* n_launches);

/* Kernel timing for roofline reports: when enabled, every pair-sum launch is bracketed by CUDA events on
 * the handle's stream.  lpmx_profile_read synchronises the stream, returns the number of pair-sum launches
 * and their summed device time (ms) and pair visits since the last read, then resets the counters. */
int lpmx_profile_enable(lpmx_handle_t h, int enable);
int lpmx_profile_read(lpmx_handle_t h, long* n_launches, double* total_ms, double* pair_visits);

/* Synchronous copy of `bytes` bytes between any two of {host, device} buffers, ordered after the work queued on
 * the handle's stream (what Kokkos::deep_copy is to the reference's update_host/update_device,
 * src/mesh/lpm_faces.hpp:151-165).  Lets host-only callers (the C++ shim's Laplacian trampoline) move data without
 * the CUDA headers. */
int lpmx_copy(lpmx_handle_t h, void* dst, const void* src, long bytes);

/* Target sharding over the GPUs of one box (the reference has no multi-device path).  This handle
 * evaluates the targets of rank `rank` of `world`: for the BVE / Incompressible2D solvers 1/world of the leaf faces and 1/world
 * of the other targets (lpmx_local_targets), for the SWE and planar solvers a contiguous range of the concatenated list
 * (vertices then faces).  Default: rank 0 of 1. */
int lpmx_set_partition(lpmx_handle_t h, int rank, int world);
/* The targets this handle's rank owns in the BVE / Incompressible2D solvers, as indices into the concatenated list (n_first
 * vertices / passive particles, then n_second faces / active particles): first its n_sources leaf faces (slice `rank` of the
 * leaves in index order), then its n_other non-sources (slice `rank` of [vertices, divided faces in index order]).
 * mask_second: HOST array of the faces' mask bytes.  idx: room for n_first + n_second entries. */
int lpmx_local_targets(lpmx_handle_t h, int n_first, int n_second, const unsigned char* mask_second, int* idx, int* n_sources,
                       int* n_other);
/* Sharded host I/O for world > 1 (BVE and Incompressible2D solvers and their in-place steppers).  Off (default): every rank
 * passes the full state and gets the full state back (replicated, as if it were alone).  On: of the HOST arrays passed to
 * set_state / *_rk?_step only the rows of lpmx_local_targets are read (area and mask are always read in full), and get_state / the
 * in-place steppers write back only those rows, without gathering the other ranks' -- each rank's host memory then holds its
 * own shard, as in any distributed-memory program.  Cuts the per-step host traffic of an 8-GPU run from 14 + 13 MB to
 * 2.9 + 1.6 MB per rank at cubed-7.  Device-pointer arguments are unaffected.  No counterpart in the reference. */
int lpmx_set_io_sharded(lpmx_handle_t h, int on);
/* NCCL bootstrap: rank 0 calls lpmx_comm_unique_id, the host program ships the 128 bytes to
 * all ranks (torch.distributed / MPI / a file), then every rank calls lpmx_comm_init. */
int lpmx_comm_unique_id(void* id128);
int lpmx_comm_init(lpmx_handle_t h, const void* id128, int rank, int world);
/* Peer exchange (opt-in, collective, after lpmx_comm_init and before the solvers are created; world <= 8, one
 * process per GPU of ONE box).  Solver slabs created afterwards are mapped into every rank with CUDA IPC and the
 * per-stage all-gather of the packed source records becomes one kernel that stores each rank's segment straight
 * into its peers over NVLink (ready / done flags with system-scope release-acquire; every wait has a deadline,
 * LPMX_PEER_TIMEOUT_S, default 600 s, after which lpmx_sync returns LPMX_ERR_COMM instead of hanging).  Returns
 * LPMX_ERR_UNSUPPORTED -- and leaves the NCCL exchange in place -- when the GPUs cannot map each other's memory.
 * lpmx_comm_init calls this itself for world <= 8 unless LPMX_PEER_EXCHANGE=0 is set in the environment.  Results are bit-identical to the
 * NCCL exchange (the same records land in the same places).  The reference has no counterpart. */
int lpmx_comm_enable_peer_exchange(lpmx_handle_t h, int enable);
/* enabled: 1 when exchanges of mapped slabs take the peer path; n_regions: slabs currently mapped. */
int lpmx_comm_peer_exchange_enabled(lpmx_handle_t h, int* enabled, int* n_regions);

/* Velocity pair sums with the source records streamed through constant banks (lpm_b200/csrc/lpmx_const_stream.cu,
 * DESIGN.md section 4.1b): sources reach the DFMAs as uniform-register operands (94-96 % of the FP64 peak issued over whole
 * evaluations against 80 % for the default kernel).  24 banks in rotation, the launches pipelined with programmatic dependent
 * launch, one CUDA graph launch per evaluation.
 * mode -1 (default): LPMX_CONST_STREAM from the environment, else AUTO = used for velocity launches from 16 384 targets where the
 * planner's modelled time beats the default kernel's by 2 % (cubed-7 on one to eight GPUs, everything larger); 0: off; 1: forced,
 * bank refills overlapped with the launches; 2: forced, refills on the compute stream.  Affects the plans made AFTER the call
 * (solvers created / states set afterwards).  The banks are __constant__ arrays of 24 modules, one set per device: the first
 * handle of a process that takes the path on a device owns them until lpmx_destroy (or mode 0); other handles on that device
 * keep the default kernel.  Same per-pair arithmetic as the default kernel; per-target sums differ by round-off. */
int lpmx_pair_sum_const_stream(lpmx_handle_t h, int mode);
/* How that path would run n_tgt targets x n_src sources on a GPU with num_sms SMs (host-only planning query): T targets per
 * thread, n_warps compute warps per CTA, ctas CTAs per bank launch covering the first n_const targets (all of them with
 * pipelined launches; with LPMX_CONST_PDL=0 whole waves, the other n_tgt - n_const go through the default kernel);
 * model_seconds / ring_seconds (may be null): the modelled duration of the evaluation on that path and on the default kernel
 * alone -- AUTO takes the path when the former is below 0.98 x the latter. */
int lpmx_const_stream_split(int num_sms, int n_tgt, int n_src, int* T, int* n_warps, int* ctas, int* n_const,
                            double* model_seconds, double* ring_seconds);
/* Bank-kernel launches issued by this handle so far (0 = every pair sum went through the default kernel): lets a caller
 * (bench.py's roofline block) name the kernel that did the work. */
int lpmx_const_stream_launch_count(lpmx_handle_t h, long* n);

/* Measured FP64 FMA throughput of this GPU in TFLOP/s (dependent-free DFMA loop on all SMs);
 * the roofline denominator that MEASURED_PEAKS.json lacks. */
int lpmx_fp64_peak_tflops(lpmx_handle_t h, double* tflops, double* ms);

/* ------------------------------------------------------------------------------------------
 * Stateless direct sums (operator level).
 *
 * Common arguments: n_tgt targets with coordinates tgt_xyz (layout/ld as above); n_src sources
 * src_xyz with vorticity src_vort[n_src], panel area src_area[n_src], mask src_mask[n_src]
 * (non-zero = divided panel, skipped as a source).  `collocated` != 0 means targets ARE the
 * sources (tgt_xyz is ignored and may be NULL, n_tgt must equal n_src) and the j == i term is
 * skipped by index.  Outputs use the same layout/ld as their target coordinates.
 * ------------------------------------------------------------------------------------------ */

/* BVEVertexVelocity (collocated=0, src/lpm_bve_sphere_kernels.hpp:179-211) and BVEFaceVelocity
 * (collocated=1, :365-394) with VelocityReduceDistinct/Collocated (:53-82, :249-274) and
 * biot_savart (src/lpm_sphere_functions.hpp:44-57).  out_vel: Real*[3]. */
int lpmx_bve_velocity(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                      const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                      const double* src_area, const unsigned char* src_mask, int n_src,
                      int collocated, double* out_vel);

/* BVEVertexStreamFn (:141-170) / BVEFaceStreamFn (:329-356) with StreamReduce* (:20-48,:218-242)
 * and greens_fn (src/lpm_sphere_functions.hpp:21-29).  out_psi: Real[n_tgt]. */
int lpmx_bve_streamfn(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                      const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                      const double* src_area, const unsigned char* src_mask, int n_src,
                      int collocated, double* out_psi);

/* BVEVertexSolve (collocated=0, src/lpm_bve_sphere_kernels.hpp:90-132) / BVEFaceSolve (collocated=1, :284-320): stream function
 * AND velocity of the same targets, which the reference computes as two nested reductions per team
 * (StreamReduce* then VelocityReduce*).  Here: ONE pass of the fused (u, psi) pair kernel, i.e. the reciprocal and the
 * logarithm of the same d = 1 - x.y per pair.  Values equal lpmx_bve_streamfn + lpmx_bve_velocity up to summation order.
 * out_psi: Real[n_tgt], out_vel: Real*[3].  (No caller inside the reference; provided for completeness of the functor set.) */
int lpmx_bve_solve(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                   const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                   const double* src_area, const unsigned char* src_mask, int n_src,
                   int collocated, double* out_psi, double* out_vel);

/* Incompressible2DPassiveSums<SphereGeometry> (collocated_targets=0,
 * src/lpm_incompressible2d_kernels.hpp:144-193) and Incompressible2DActiveSums (collocated
 * targets, :201-246; the self term is skipped only when |eps| < DBL_EPSILON, :235), with
 * kernel_vals (:35-47) and Incompressible2DReducer (:92-136).  out_psi may be NULL (velocity
 * only). */
int lpmx_ic2d_sums(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld, int n_tgt,
                   const double* src_xyz, int src_layout, long src_ld, const double* src_vort,
                   const double* src_area, const unsigned char* src_mask, int n_src, double eps,
                   int targets_are_sources, double* out_vel, double* out_psi);

/* SphereVertexSums (targets_are_sources=0, src/lpm_swe_kernels.hpp:723-780) and SphereFaceSums
 * (:877-930) with SphereSweDirectSumReducer (:578-619) and sphere_swe_velocity_sums (:334-362).
 * out_vel (Real*[3]) is written only if do_velocity != 0; out_ddot[n_tgt] always.
 * out_grad (optional, may be NULL): the 9 accumulated gradient sums per target, row-major
 * [n_tgt][9] regardless of layout (a debugging/validation aid; the reference does not expose
 * them). */
int lpmx_swe_sphere_sums(lpmx_handle_t h, const double* tgt_xyz, int tgt_layout, long tgt_ld,
                         int n_tgt, const double* src_xyz, int src_layout, long src_ld,
                         const double* src_vort, const double* src_div, const double* src_area,
                         const unsigned char* src_mask, int n_src, double eps,
                         int targets_are_sources, int do_velocity, double* out_vel,
                         double* out_ddot, double* out_grad);

/* ------------------------------------------------------------------------------------------
 * Time steppers (stepper level).
 * ------------------------------------------------------------------------------------------ */

/* BVERK4::advance_timestep(vx, vzeta, vvel, fx, fzeta, fvel, fa, fm) (src/lpm_bve_rk4.hpp:47-49,
 * src/lpm_bve_rk4_impl.hpp:63-167), repeated n_steps times, IN PLACE on the caller's views.
 * On entry vvel/fvel must hold the velocity of the current state (as after
 * BVESphere::init_velocity, src/lpm_bve_sphere_impl.hpp:181-207); on exit they hold the
 * velocity of the new state.  Replicates the reference's face-vorticity update as coded
 * (facevort4 in the facevort3 slot, :155-157). */
int lpmx_bve_rk4_step(lpmx_handle_t h, double dt, double Omega, int n_verts, double* vert_xyz,
                      double* vert_vort, double* vert_vel, int n_faces, double* face_xyz,
                      double* face_vort, double* face_vel, const double* face_area,
                      const unsigned char* face_mask, int layout, long vert_ld, long face_ld,
                      int n_steps);

/* Persistent-state variant: state lives in HBM across calls (what BVERK4 + BVESphere's device
 * views do in the reference); set/get move it.  All arrays as in lpmx_bve_rk4_step. */
int lpmx_bve_solver_create(lpmx_handle_t h, int n_verts, int n_faces, lpmx_bve_solver_t* s);
int lpmx_bve_solver_destroy(lpmx_bve_solver_t s);
int lpmx_bve_solver_set_state(lpmx_bve_solver_t s, const double* vert_xyz, const double* vert_vort,
                              const double* vert_vel, const double* face_xyz,
                              const double* face_vort, const double* face_vel,
                              const double* face_area, const unsigned char* face_mask, int layout,
                              long vert_ld, long face_ld);
/* any output pointer may be NULL */
int lpmx_bve_solver_get_state(lpmx_bve_solver_t s, double* vert_xyz, double* vert_vort,
                              double* vert_vel, double* face_xyz, double* face_vort,
                              double* face_vel, int layout, long vert_ld, long face_ld);
/* BVESphere::init_velocity (src/lpm_bve_sphere_impl.hpp:181-207) on the resident state */
int lpmx_bve_solver_init_velocity(lpmx_bve_solver_t s);
/* BVESphere::init_stream_fn (:209-231): psi at vertices then faces, into out arrays (host or
 * device, either may be NULL) */
int lpmx_bve_solver_stream_fn(lpmx_bve_solver_t s, double* vert_psi, double* face_psi);
int lpmx_bve_solver_advance(lpmx_bve_solver_t s, double dt, double Omega, int n_steps);
/* pair interactions evaluated by one velocity evaluation of this solver on this rank
 * ((n_v + n_f) * n_leaf - n_leaf over all ranks) */
int lpmx_bve_solver_interactions_per_eval(lpmx_bve_solver_t s, double* local, double* global);

/* Incompressible2DRK2::advance_timestep_impl (src/lpm_incompressible2d_rk2_impl.hpp:75-172)
 * for SphereGeometry, n_steps times, in place.  passive = vertices, active = faces.
 * Omega is CoriolisSphere::Omega (src/lpm_coriolis.hpp:154-195).  On entry the velocities must
 * hold the current state's velocity (Incompressible2D::init_direct_sums,
 * src/lpm_incompressible2d_impl.hpp:235-254); on exit velocity and stream function are those
 * of the new state.  With n_steps > 1 the stream function is evaluated for the last step only: the reference overwrites it
 * at every step, so the intermediate values are never observable from a multi-step call. */
int lpmx_ic2d_rk2_step(lpmx_handle_t h, double dt, double Omega, double eps, int n_passive,
                       double* passive_xyz, double* passive_vort, double* passive_vel,
                       double* passive_psi, int n_active, double* active_xyz, double* active_vort,
                       double* active_vel, double* active_psi, const double* active_area,
                       const unsigned char* active_mask, int layout, long passive_ld,
                       long active_ld, int n_steps);

int lpmx_ic2d_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps,
                            lpmx_ic2d_solver_t* s);
int lpmx_ic2d_solver_destroy(lpmx_ic2d_solver_t s);
int lpmx_ic2d_solver_set_state(lpmx_ic2d_solver_t s, const double* passive_xyz,
                               const double* passive_vort, const double* passive_vel,
                               const double* active_xyz, const double* active_vort,
                               const double* active_vel, const double* active_area,
                               const unsigned char* active_mask, int layout, long passive_ld,
                               long active_ld);
int lpmx_ic2d_solver_get_state(lpmx_ic2d_solver_t s, double* passive_xyz, double* passive_vort,
                               double* passive_vel, double* passive_psi, double* active_xyz,
                               double* active_vort, double* active_vel, double* active_psi,
                               int layout, long passive_ld, long active_ld);
/* Incompressible2D::init_direct_sums on the resident state */
int lpmx_ic2d_solver_init_direct_sums(lpmx_ic2d_solver_t s);
/* n_steps x Incompressible2DRK2::advance_timestep_impl on the resident state.  The stream function of the new state is an
 * output of a step that no later step reads (the predictor's psi is overwritten by the corrector's, and the corrector's by the
 * next step's: src/lpm_incompressible2d_rk2_impl.hpp:113-124,157-170), and evaluating it costs more than the velocity.  It is
 * therefore LAZY: advance() fuses psi into its last evaluation only if psi was read (get_state with a psi pointer) since the
 * previous advance; otherwise it runs velocity-only evaluations and the next reader of psi triggers one psi-only pass over the
 * retained state.  Values are the same sums either way (round-off of the summation order aside).  All ranks of a sharded
 * solver must make the same get_state calls (they already must: get_state gathers). */
int lpmx_ic2d_solver_advance(lpmx_ic2d_solver_t s, double dt, double Omega, int n_steps);
/* Override the laziness for the next advance: demand_next != 0 fuses psi into its final evaluation, 0 leaves it stale. */
int lpmx_ic2d_solver_lazy_stream_fn(lpmx_ic2d_solver_t s, int demand_next);

/* ------------------------------------------------------------------------------------------
 * Per-step diagnostics of the callers (the O(N) tail of every example's time loop).
 * ------------------------------------------------------------------------------------------ */

/* Incompressible2D::total_vorticity / total_kinetic_energy / total_enstrophy
 * (src/lpm_incompressible2d_impl.hpp:91-137): sums over the unmasked active particles of zeta A, |u|^2 A / 2,
 * zeta^2 A / 2.  Host or device pointers; any output may be NULL. */
int lpmx_ic2d_totals(lpmx_handle_t h, int n_active, const double* active_vort, const double* active_vel, int layout,
                     long active_ld, const double* active_area, const unsigned char* active_mask,
                     double* total_vorticity, double* total_kinetic_energy, double* total_enstrophy);
/* the same on the resident state of a solver (no transfer of the fields) */
int lpmx_ic2d_solver_totals(lpmx_ic2d_solver_t s, double* total_vorticity, double* total_kinetic_energy,
                            double* total_enstrophy);

/* ErrNorms(err, exact, weight) (src/lpm_error.hpp:81-131, ReduceErrorFtor src/lpm_error_impl.hpp:59-108):
 * l1 = sum |e| w / sum |x| w, l2 = sqrt(sum e^2 w / sum x^2 w), linf = max |e| / max |x|; ndim = 1 for scalar
 * views, 3 for Real*[3] views (|.| = Euclidean magnitude of a row; layout/ld as everywhere). */
int lpmx_err_norms(lpmx_handle_t h, int n, int ndim, const double* err, const double* exact, int layout, long ld,
                   const double* weight, double* l1, double* l2, double* linf);

/* ComputeFTLE<SeedType> for quadrilateral faces (src/mesh/lpm_ftle.hpp:15-319; launched once per step by
 * examples/sphere_rh54.cpp:308-316, sphere_gaussian_vortex.cpp:252-260, plane_colliding_dipoles.cpp:269-277) and
 * get_max_ftle (:327-338).  geom selects the SphereGeometry (Real*[3]) or PlaneGeometry (Real*[2]) branch;
 * `layout`/`*_ld` describe the four coordinate views, `verts_layout` the Index*[4] view faces.verts
 * (LPMX_LAYOUT_RIGHT: v[f*4+k], LPMX_LAYOUT_LEFT: v[k*n_faces+f]).  As in the reference: ftle(f) = log(lambda_1) of
 * the elementwise product F_ij F_ji (no division by 2t); entries of divided faces (mask != 0) are not written; on
 * the sphere face_phys(f,:) of every leaf is normalised IN PLACE (:98), so face_phys is an in/out argument there;
 * triangular faces are a static_assert in the reference (:19-20) and are not offered.  Host or device pointers;
 * max_ftle may be NULL. */
#define LPMX_GEOM_SPHERE 0
#define LPMX_GEOM_PLANE 1
int lpmx_ftle(lpmx_handle_t h, int geom, int n_verts, const double* vert_phys, const double* vert_ref, int layout,
              long vert_ld, int n_faces, double* face_phys, const double* face_ref, long face_ld, const int* face_verts,
              int verts_layout, const unsigned char* face_mask, double* ftle, double* max_ftle);

/* ------------------------------------------------------------------------------------------
 * Spherical shallow water: SWE<Seed> fields + SWERK2 (src/lpm_swe.hpp:29-88, src/lpm_swe_rk2.hpp:15-39).
 * ------------------------------------------------------------------------------------------ */

/* The SWE<Seed> fields the stepper reads and writes.  passive = vertices, active = faces.  Array lengths are
 * n_passive / n_active; xyz and vel are Real*[3] in the call's layout.  (Reference names: mesh.vertices.phys_crds,
 * rel_vort_passive, div_passive, depth_passive, surf_passive, bottom_passive, velocity_passive,
 * double_dot_passive, surf_lap_passive; mesh.faces.phys_crds, rel_vort_active, div_active, mesh.faces.area,
 * mass_active, depth_active, surf_active, bottom_active, velocity_active, double_dot_active, surf_lap_active,
 * mesh.faces.mask.) */
typedef struct lpmx_swe_passive_s {
  double *xyz, *vort, *div, *depth, *surf, *bottom, *vel, *ddot, *laps;
} lpmx_swe_passive_t;
typedef struct lpmx_swe_active_s {
  double *xyz, *vort, *div, *area, *mass, *depth, *surf, *bottom, *vel, *ddot, *laps;
  const unsigned char* mask;
} lpmx_swe_active_t;

/* Surface-Laplacian provider.  SWERK2 evaluates the Laplacian of the surface height with Compadre GMLS on a
 * host kd-tree (src/lpm_swe_rk2_impl.hpp:134-154 for the predictor state = stage 1, :233-252 for the new state =
 * stage 2); that third-party step is outside this library and is supplied by the caller.  All pointers are DEVICE
 * pointers; xyz arrays are LPMX_LAYOUT_LEFT with leading dimension xyz_ld.  The provider must fill passive_laps
 * and active_laps, ordering its work on `cuda_stream` (or synchronising that stream before host access), and
 * return 0.  A NULL provider leaves the Laplacian arrays as they are (see lpmx_swe_solver_set_laplacian). */
typedef int (*lpmx_swe_laplacian_fn)(void* user, int stage, void* cuda_stream, int n_passive,
                                     const double* passive_xyz, const double* passive_surf, double* passive_laps,
                                     int n_active, const double* active_xyz, const double* active_surf,
                                     const unsigned char* active_mask, double* active_laps, long xyz_ld);

/* SWERK2<Seed, ZeroFunctor>::advance_timestep_impl (src/lpm_swe_rk2_impl.hpp:80-258) for SphereGeometry, n_steps
 * times, IN PLACE on the caller's arrays (host or device pointers).  On entry vel, ddot and laps must belong to
 * the current state (SWE::init_direct_sums, src/lpm_swe_impl.hpp:401-445, and the SWERK2 constructor's Laplacian,
 * rk2_impl.hpp:56-77); every field of both structs is required except surf/bottom/depth(active)/laps/div, which
 * default to zero when NULL.  `eps` is the kernel smoothing parameter (never initialised by the reference on
 * the sphere, SURVEY.md quirk C-iii: pass it explicitly).  Bottom topography is ZeroFunctor. */
int lpmx_swe_rk2_step(lpmx_handle_t h, double dt, double Omega, double g, double eps, int n_passive,
                      const lpmx_swe_passive_t* passive, int n_active, const lpmx_swe_active_t* active, int layout,
                      long passive_ld, long active_ld, lpmx_swe_laplacian_fn laplacian, void* user, int n_steps);

/* Persistent-state variant */
int lpmx_swe_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps, lpmx_swe_solver_t* s);
int lpmx_swe_solver_destroy(lpmx_swe_solver_t s);
int lpmx_swe_solver_set_state(lpmx_swe_solver_t s, const lpmx_swe_passive_t* passive, const lpmx_swe_active_t* active,
                              int layout, long passive_ld, long active_ld);
/* any pointer inside the structs may be NULL (= not wanted) */
int lpmx_swe_solver_get_state(lpmx_swe_solver_t s, const lpmx_swe_passive_t* passive, const lpmx_swe_active_t* active,
                              int layout, long passive_ld, long active_ld);
/* overwrite the resident surface-Laplacian arrays (host or device pointers; either may be NULL) */
int lpmx_swe_solver_set_laplacian(lpmx_swe_solver_t s, const double* passive_laps, const double* active_laps);
/* SWE::init_direct_sums(do_velocity) on the resident state (physical == Lagrangian coordinates at t = 0) */
int lpmx_swe_solver_init_direct_sums(lpmx_swe_solver_t s, int do_velocity);
int lpmx_swe_solver_advance(lpmx_swe_solver_t s, double dt, double Omega, double g, lpmx_swe_laplacian_fn laplacian,
                            void* user, int n_steps);

/* ------------------------------------------------------------------------------------------
 * The steps either side of the spherical SWE direct sums, kept on the device: the reference
 * gathers vertices + leaf faces, copies them to the host, builds Compadre neighbourhoods on a kd-tree and evaluates a
 * GMLS surface Laplacian every RK stage (src/lpm_swe_rk2_impl.hpp:56-77,134-154,233-252).
 * ------------------------------------------------------------------------------------------ */

/* GatherMeshData<Seed> (src/mesh/lpm_gather_mesh_data.hpp:24-117, functors _impl.hpp:14-66): rows [0, n_verts) of
 * `gathered` are the vertex rows, row n_verts + leaf_idx(f) is leaf face f (faces.leaf_idx = exclusive scan of
 * !mask); divided faces are dropped.  n_comp = 1 (scalar fields) or 2 / 3 (coordinates, vector fields); all three
 * arrays share `layout`.  Call with gathered == NULL to obtain *n_gathered = n_verts + n_leaves only. */
int lpmx_gather_mesh_data(lpmx_handle_t h, int n_comp, int layout, int n_verts, const double* vert_data, long vert_ld,
                          int n_faces, const double* face_data, long face_ld, const unsigned char* face_mask,
                          double* gathered, long gathered_ld, int* n_gathered);
/* ScatterMeshData<Seed>::scatter_fields (src/mesh/lpm_scatter_mesh_data_impl.hpp): the inverse; rows of divided
 * faces in face_data are left as they are. */
int lpmx_scatter_mesh_data(lpmx_handle_t h, int n_comp, int layout, const double* gathered, long gathered_ld, int n_verts,
                           double* vert_data, long vert_ld, int n_faces, double* face_data, long face_ld,
                           const unsigned char* face_mask);

/* gmls::Params (src/lpm_compadre.hpp:23-60), same members and defaults */
typedef struct lpmx_gmls_params_s {
  double eps_multiplier;      /* window radius = eps_multiplier x distance to the min_neighbors-th nearest point */
  int samples_order;          /* Taylor order of the data reconstruction, 2..4 */
  int manifold_order;         /* Taylor order of the manifold (height) reconstruction, 1..4 */
  double samples_weight_pwr;  /* p of the Power weight (1 - r/eps)^p */
  double manifold_weight_pwr; /* must equal samples_weight_pwr (both are 2 in every gmls::Params constructor) */
  int ambient_dim;            /* 3 */
  int topo_dim;               /* 2 */
  int min_neighbors;          /* Compadre::GMLS::getNP(samples_order, topo_dim) = (order+1)(order+2)/2 by default */
} lpmx_gmls_params_t;
/* gmls::Params(order, 3) */
int lpmx_gmls_params_init(lpmx_gmls_params_t* params, int order);

/* Surface Laplacian (Laplace-Beltrami) of the samples f at n collocated source/target points on a sphere:
 * gmls::Neighborhoods(crds, params) + gmls::sphere_scalar_gmls(crds, crds, neighbors, params,
 * {LaplacianOfScalarPointEvaluation}) + Evaluator::applyAlphasToDataAllComponentsAllTargetSites
 * (src/lpm_swe_rk2_impl.hpp:66-75).  Compadre 1.6.2 is not part of the reference tree: this is its published algorithm
 * (lpm_b200/csrc/lpmx_gmls_core.h), validated analytically, NOT pinned value-for-value.
 * Optional outputs: the window radius and the neighbour count of every point (Neighborhoods::neighborhood_radii,
 * neighbor_lists(i, 0)).  A point whose least-squares system is rank deficient gets NaN.  Host or device pointers. */
int lpmx_gmls_sphere_laplacian(lpmx_handle_t h, const lpmx_gmls_params_t* params, int n, const double* xyz, int layout,
                               long ld, const double* f, double* laplacian, double* window_radius, int* n_neighbors);

/* Scalar point evaluation of n_fields source fields at n_tgt target points on the sphere: gmls::Neighborhoods(src, tgt,
 * params) + ScalarPointEvaluation / PointSample, the interpolation step of CompadreRemesh (interpolate_lag_crds and
 * uniform_direct_remesh, src/mesh/lpm_compadre_remesh_impl.hpp:136-210; driven from examples/sphere_rh54.cpp:257-300).
 * src_fields / tgt_fields are arrays of n_fields pointers (each n_src / n_tgt doubles; host or device, also the pointer
 * arrays themselves live on the host).  samples_order 1..4; the manifold reconstruction does not enter a point
 * evaluation.  Same status as the Laplacian: Compadre's published algorithm, validated analytically. */
int lpmx_gmls_sphere_interpolate(lpmx_handle_t h, const lpmx_gmls_params_t* params, int n_src, const double* src_xyz,
                                 int src_layout, long src_ld, int n_fields, const double* const* src_fields, int n_tgt,
                                 const double* tgt_xyz, int tgt_layout, long tgt_ld, double* const* tgt_fields);

/* Built-in surface-Laplacian provider for lpmx_swe_rk2_step / lpmx_swe_solver_advance: pass
 * `lpmx_gmls_swe_laplacian` as the lpmx_swe_laplacian_fn and a lpmx_gmls_provider_t* as `user`.  It performs the
 * reference's gather -> neighbourhoods -> GMLS -> scatter sequence without leaving the device. */
typedef struct lpmx_gmls_provider_s {
  lpmx_handle_t handle;
  lpmx_gmls_params_t params;
} lpmx_gmls_provider_t;
int lpmx_gmls_swe_laplacian(void* user, int stage, void* cuda_stream, int n_passive, const double* passive_xyz,
                            const double* passive_surf, double* passive_laps, int n_active, const double* active_xyz,
                            const double* active_surf, const unsigned char* active_mask, double* active_laps, long xyz_ld);

/* ------------------------------------------------------------------------------------------
 * Adaptive refinement flags (src/mesh/lpm_refinement_flags.hpp, src/mesh/lpm_refinement.hpp): O(N) streaming kernels
 * over the faces.  A flag functor turns flags[i] on where its criterion holds on an undivided face and never turns a
 * flag off; Refinement<Seed>::iterate (lpm_refinement.hpp:28-41) clears the flags, runs the functor over
 * [start, end) and counts.  NeighborsFlag (:30-53, marked "likely won't work on device", used by no driver) is not provided.
 * ------------------------------------------------------------------------------------------ */
#define LPMX_FLAG_SCALAR_MAX 0         /* ScalarMaxFlag (:132-183):        |f_i| > tol                                  */
#define LPMX_FLAG_SCALAR_INTEGRAL 1    /* ScalarIntegralFlag (:185-229):   |f_i| A_i > tol                              */
#define LPMX_FLAG_SCALAR_VARIATION 2   /* ScalarVariationFlag (:231-310):  max - min over {f_i, f at the face's vertices} > tol */
#define LPMX_FLAG_FLOW_MAP_VARIATION 3 /* FlowMapVariationFlag (:55-130):  sum_k (max - min over the face's vertices of lag_k) > tol */

typedef struct lpmx_flag_desc_s {
  int kind;                  /* LPMX_FLAG_*                                                                      */
  int n_faces, n_verts;      /* mesh.n_faces_host(), mesh.n_vertices_host()                                      */
  int n_face_verts;          /* 3 or 4                                                                           */
  const double* face_vals;   /* [n_faces]  kinds 0-2                                                             */
  const double* area;        /* [n_faces]  kind 1                                                                */
  const double* vert_vals;   /* [n_verts]  kind 2                                                                */
  const int* face_verts;     /* [n_faces][n_face_verts] row-major, kinds 2-3                                     */
  const double* vert_lag;    /* Real*[ndim] vertex Lagrangian coordinates, kind 3                                */
  int ndim, layout;          /* 3 (sphere) or 2 (plane); LPMX_LAYOUT_* of vert_lag                               */
  long ld;
  const unsigned char* mask; /* [n_faces]  faces.mask                                                            */
  double tol;                /* absolute tolerance used by lpmx_refine_flag                                      */
} lpmx_flag_desc_t;

/* The reduction of <Flag>::set_tol_from_relative_value(): the maximum the relative tolerance multiplies (kind 0:
 * max |f_i| over ALL faces; 1: max |f_i| A_i over all faces; 2, 3: the variation over undivided faces).  As in the
 * reference the caller then sets tol = relative_tol * max_value.  Host or device pointers. */
int lpmx_refine_flag_max(lpmx_handle_t h, const lpmx_flag_desc_t* flag, double* max_value);

/* Run the flag functor over faces [start, end): flags[i] |= criterion(i) for undivided faces.  flags: n_faces bytes,
 * in/out (host or device); clear_first != 0 zeroes all n_faces entries first (Refinement::iterate).  *count = number of
 * non-zero flags in [start, end) afterwards (nullable). */
int lpmx_refine_flag(lpmx_handle_t h, const lpmx_flag_desc_t* flag, int start, int end, int clear_first,
                     unsigned char* flags, int* count);

/* ------------------------------------------------------------------------------------------
 * Planar problems (PlaneGeometry): Real*[2] views.  LPMX_LAYOUT_RIGHT is x[i*2+k], LPMX_LAYOUT_LEFT is x[k*ld+i].
 * Coriolis is CoriolisBetaPlane(f0, beta) (src/lpm_coriolis.hpp:93-148): f = f0 + beta y.
 * ------------------------------------------------------------------------------------------ */

/* bottom topography functors of src/lpm_surface_gallery.hpp usable in the plane */
#define LPMX_TOPO_ZERO 0                     /* ZeroFunctor (:92-102) */
#define LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN 1 /* PlanarGaussianMountain (:41-61): 0.8 exp(-5 |x|^2) */

/* Incompressible2DPassiveSums<PlaneGeometry> (targets_are_sources=0, src/lpm_incompressible2d_kernels.hpp:144-193)
 * and Incompressible2DActiveSums<PlaneGeometry> (:201-246; the self term is skipped only when |eps| < zero_tol),
 * with Incompressible2DKernels<PlaneGeometry>::kernel_vals (:55-85).  out_vel: Real*[2] in the targets' layout;
 * out_psi may be NULL. */
int lpmx_ic2d_plane_sums(lpmx_handle_t h, const double* tgt_xy, int tgt_layout, long tgt_ld, int n_tgt,
                         const double* src_xy, int src_layout, long src_ld, const double* src_vort,
                         const double* src_area, const unsigned char* src_mask, int n_src, double eps,
                         int targets_are_sources, double* out_vel, double* out_psi);

/* Incompressible2DRK2<Seed>::advance_timestep_impl (src/lpm_incompressible2d_rk2_impl.hpp:75-172) for PlaneGeometry
 * (examples/plane_colliding_dipoles.cpp:196), n_steps times, in place.  Arguments as lpmx_ic2d_rk2_step. */
int lpmx_ic2d_plane_rk2_step(lpmx_handle_t h, double dt, double f0, double beta, double eps, int n_passive,
                             double* passive_xy, double* passive_vort, double* passive_vel, double* passive_psi,
                             int n_active, double* active_xy, double* active_vort, double* active_vel,
                             double* active_psi, const double* active_area, const unsigned char* active_mask,
                             int layout, long passive_ld, long active_ld, int n_steps);

/* The planar SWE<Seed> fields (src/lpm_swe.hpp:29-88).  passive = vertices, active = faces; xy and vel are Real*[2].
 * Reference names: mesh.vertices.phys_crds, rel_vort_passive, div_passive, depth_passive, surf_passive,
 * bottom_passive, velocity_passive, double_dot_passive, du1dx1_passive .. du2dx2_passive, surf_lap_passive,
 * stream_fn_passive, potential_passive; the active ones likewise plus mesh.faces.area, mass_active, mesh.faces.mask. */
typedef struct lpmx_plane_swe_passive_s {
  double *xy, *vort, *div, *depth, *surf, *bottom, *vel, *ddot, *du1dx1, *du1dx2, *du2dx1, *du2dx2, *laps, *psi, *phi;
} lpmx_plane_swe_passive_t;
typedef struct lpmx_plane_swe_active_s {
  double *xy, *vort, *div, *area, *mass, *depth, *surf, *bottom, *vel, *ddot, *du1dx1, *du1dx2, *du2dx1, *du2dx2, *laps,
      *psi, *phi;
  const unsigned char* mask;
} lpmx_plane_swe_active_t;

/* The nine direct-sum outputs of PlanarSWEVertexSums / PlanarSWEFaceSums (any pointer may be NULL) */
typedef struct lpmx_plane_swe_sums_s {
  double *vel, *ddot, *du1dx1, *du1dx2, *du2dx1, *du2dx2, *laps, *psi, *phi;
} lpmx_plane_swe_sums_t;

/* PlanarSWEVertexSums (targets_are_sources=0, src/lpm_swe_kernels.hpp:626-716) and PlanarSWEFaceSums (:787-870;
 * the self term is skipped unless eps > 0) with PlanarSwePseDirectSumReducer (:522-571), planar_swe_sums_rhs_pse
 * (:393-445) and pse::BivariateOrder8::laplacian (src/lpm_pse.hpp:66-73).  tgt_surf/src_surf are the surface heights
 * the PSE Laplacian differences.  out->vel is written only if do_velocity != 0. */
int lpmx_swe_plane_sums(lpmx_handle_t h, const double* tgt_xy, int tgt_layout, long tgt_ld, const double* tgt_surf,
                        int n_tgt, const double* src_xy, int src_layout, long src_ld, const double* src_vort,
                        const double* src_div, const double* src_area, const unsigned char* src_mask,
                        const double* src_surf, int n_src, double eps, double pse_eps, int targets_are_sources,
                        int do_velocity, const lpmx_plane_swe_sums_t* out);

/* SWERK4<Seed, Topo>::advance_timestep (src/lpm_swe_rk4.hpp:88-96, src/lpm_swe_rk4_impl.hpp:203-445) for
 * PlaneGeometry (examples/plane_gravity_wave.cpp:171-176), n_steps times, IN PLACE on the caller's arrays (host or
 * device pointers).  On entry vel, ddot and laps must belong to the current state (SWE::init_direct_sums,
 * src/lpm_swe_impl.hpp:401-422).  Required: xy, vort, depth (passive) / area, mass, mask (active), vel, ddot, laps;
 * the rest default to zero when NULL on input and are skipped on output.  As coded in the reference, the
 * fourth-stage position increment is never assigned (x4 = 0), so positions advance with
 * x += (x1 + 0)/6 + (x2 + x3)/3; this is replicated. */
int lpmx_swe_plane_rk4_step(lpmx_handle_t h, double dt, double f0, double beta, double g, double eps, double pse_eps,
                            int topo, int n_passive, const lpmx_plane_swe_passive_t* passive, int n_active,
                            const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld,
                            int n_steps);

/* Persistent-state variant (state resident in HBM across calls) */
int lpmx_plane_swe_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps, double pse_eps, int topo,
                                 lpmx_plane_solver_t* s);
int lpmx_plane_swe_solver_destroy(lpmx_plane_solver_t s);
int lpmx_plane_swe_solver_set_state(lpmx_plane_solver_t s, const lpmx_plane_swe_passive_t* passive,
                                    const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld);
/* any pointer inside the structs may be NULL (= not wanted) */
int lpmx_plane_swe_solver_get_state(lpmx_plane_solver_t s, const lpmx_plane_swe_passive_t* passive,
                                    const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld);
/* SWE::init_direct_sums(do_velocity) on the resident state */
int lpmx_plane_swe_solver_init_direct_sums(lpmx_plane_solver_t s, int do_velocity);
int lpmx_plane_swe_solver_advance(lpmx_plane_solver_t s, double dt, double f0, double beta, double g, int n_steps);
/* pair interactions evaluated by one direct-sum evaluation of this solver: local (this rank) and global */
int lpmx_plane_swe_solver_interactions_per_eval(lpmx_plane_solver_t s, double* local, double* global);

#ifdef __cplusplus
}
#endif
#endif /* LPMX_H */
