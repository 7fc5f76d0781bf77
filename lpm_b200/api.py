"""Host-side mirror of the reference's operator / stepper interface over the C ABI.

The reference is C++ (its API surface is mirrored in C++ under include/lpm/); this module is the
thin Python binding the parity tests and bench.py drive.  Names follow the reference:
PolyMesh2d / MeshSeed (src/mesh/lpm_polymesh2d.hpp), BVEVertexVelocity ... (src/lpm_bve_sphere_kernels.hpp),
BVERK4 (src/lpm_bve_rk4.hpp), Incompressible2DRK2 (src/lpm_incompressible2d_rk2.hpp),
SphereVertexSums / SphereFaceSums (src/lpm_swe_kernels.hpp).

Arrays may be numpy arrays (host: staged inside the call) or torch CUDA tensors (device: used in
place).  Everything is float64 / int32 / uint8 as in the reference (Real / Index / bool).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (LAYOUT_LEFT, LAYOUT_RIGHT, SEED_CUBED_SPHERE, SEED_ICOS_TRI_SPHERE, SEED_QUAD_RECT, SEED_TRI_HEX,  # noqa: F401
                   LpmxError)

_MESH_ARRAYS = {
    # width -1: vertices per face; -2: Geo::ndim (3 on the sphere, 2 in the plane; the planar arrays keep the names *_xyz)
    "vert_xyz": (0, -2), "vert_lag_xyz": (1, -2), "vert_crd_inds": (2, 1),
    "edge_origs": (3, 1), "edge_dests": (4, 1), "edge_lefts": (5, 1), "edge_rights": (6, 1),
    "edge_parents": (7, 1), "edge_kids": (8, 2),
    "face_xyz": (9, -2), "face_lag_xyz": (10, -2), "face_area": (11, 1), "face_mask": (12, 1),
    "face_verts": (13, -1), "face_edges": (14, -1), "face_crd_inds": (15, 1), "face_parent": (16, 1),
    "face_kids": (17, 4), "face_level": (18, 1), "face_leaf_idx": (19, 1),
}

SEEDS = {"icos": SEED_ICOS_TRI_SPHERE, "icostri_sphere": SEED_ICOS_TRI_SPHERE,
         "cubed": SEED_CUBED_SPHERE, "cubed_sphere": SEED_CUBED_SPHERE,
         "quad_rect": SEED_QUAD_RECT, "tri_hex": SEED_TRI_HEX}


def _seed_id(seed):
    if isinstance(seed, str):
        return SEEDS[seed]
    return int(seed)


def max_allocations(seed, depth):
    """MeshSeed<Seed>::set_max_allocations (src/mesh/lpm_mesh_seed.cpp:266-279)."""
    L = _lib.lib()
    nv, ne, nf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = L.lpmx_mesh_max_allocations(_seed_id(seed), depth, nv, ne, nf)
    if rc:
        raise LpmxError(rc, "lpmx_mesh_max_allocations")
    return nv.value, ne.value, nf.value


class PolyMesh2d:
    """Tree mesh on the sphere: PolyMesh2d<Seed>::tree_init, and divide_flagged_faces for adaptive refinement
    (host; no GPU needed).  amr_buffer / amr_limit as in PolyMeshParameters (src/mesh/lpm_polymesh2d.hpp:32-72):
    nmaxfaces is sized for depth + amr_buffer, a face may be refined amr_limit times beyond `depth`."""

    AMR_DIVIDED_ALL, AMR_NO_SPACE, AMR_LIMIT_REACHED = 0, 1, 2

    def __init__(self, seed, depth, radius=1.0, amr_buffer=0, amr_limit=0):
        L = _lib.lib()
        self.seed = _seed_id(seed)
        self.ndim = 2 if self.seed in (SEED_QUAD_RECT, SEED_TRI_HEX) else 3
        self.depth = depth
        self.amr_buffer, self.amr_limit = amr_buffer, amr_limit
        self.nmaxverts, self.nmaxedges, self.nmaxfaces = max_allocations(seed, depth + amr_buffer)
        m = ctypes.c_void_p()
        rc = L.lpmx_mesh_create(self.seed, depth, float(radius), ctypes.byref(m))
        if rc:
            raise LpmxError(rc, "lpmx_mesh_create")
        self._m = m
        self._fetch()

    def __del__(self):
        m, self._m = getattr(self, "_m", None), None
        if m:
            try:
                _lib.lib().lpmx_mesh_destroy(m)
            except Exception:
                pass

    def _fetch(self):
        L, m = _lib.lib(), self._m
        s = [ctypes.c_int() for _ in range(6)]
        L.lpmx_mesh_sizes(m, *s)
        (self.n_verts, self.n_edges, self.n_faces, self.n_face_leaves, self.n_edge_leaves,
         self.n_face_verts) = [x.value for x in s]
        for name, (aid, width) in _MESH_ARRAYS.items():
            p, n, kind = ctypes.c_void_p(), ctypes.c_long(), ctypes.c_int()
            rc = L.lpmx_mesh_array(m, aid, ctypes.byref(p), ctypes.byref(n), ctypes.byref(kind))
            if rc:
                raise LpmxError(rc, "lpmx_mesh_array")
            ctype = {0: ctypes.c_int, 1: ctypes.c_double, 2: ctypes.c_ubyte}[kind.value]
            if n.value == 0:
                arr = np.zeros(0, dtype=np.dtype(ctype))
            else:
                arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n.value,)).copy()
            w = self.n_face_verts if width == -1 else (self.ndim if width == -2 else width)
            if w > 1:
                arr = arr.reshape(-1, w)
            setattr(self, name, arr)

    def _push_coordinates(self):
        """Hand this object's coordinate arrays (which the caller may have advected) to the generator."""
        L = _lib.lib()
        for name in ("vert_xyz", "vert_lag_xyz", "face_xyz", "face_lag_xyz"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            rc = L.lpmx_mesh_update_array(self._m, _MESH_ARRAYS[name][0], a.ctypes.data, a.size)
            if rc:
                raise LpmxError(rc, "lpmx_mesh_update_array")

    def divide_flagged_faces(self, flags, push_coordinates=True):
        """PolyMesh2d::divide_flagged_faces (src/mesh/lpm_polymesh2d_impl.hpp:124-173).  The coordinate arrays of this object
        (which the caller may have advected) are handed to the generator first, the arrays are re-read afterwards.
        Returns (n_divided, outcome) with outcome one of AMR_DIVIDED_ALL / AMR_NO_SPACE / AMR_LIMIT_REACHED."""
        L = _lib.lib()
        if push_coordinates:
            self._push_coordinates()
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        nd, oc = ctypes.c_int(), ctypes.c_int()
        rc = L.lpmx_mesh_divide_flagged_faces(self._m, f.ctypes.data, f.size, self.nmaxfaces, self.depth + self.amr_limit,
                                              ctypes.byref(nd), ctypes.byref(oc))
        if rc:
            raise LpmxError(rc, "lpmx_mesh_divide_flagged_faces")
        self._fetch()
        return nd.value, oc.value

    # ---- mesh queries (src/mesh/lpm_polymesh2d.hpp:262-552), host code like the mesh ----
    def _index_list(self, fn, idx, name):
        cap = 64
        while True:
            buf = np.empty(cap, dtype=np.int32)
            n = ctypes.c_int()
            rc = fn(self._m, int(idx), buf.ctypes.data, cap, ctypes.byref(n))
            if rc:
                raise LpmxError(rc, name)
            if n.value <= cap:
                return buf[:n.value].copy()
            cap = n.value

    def get_leaf_edges_from_parent(self, parent_edge):
        return self._index_list(_lib.lib().lpmx_mesh_leaf_edges_from_parent, parent_edge, "lpmx_mesh_leaf_edges_from_parent")

    def ccw_edges_around_face(self, face):
        return self._index_list(_lib.lib().lpmx_mesh_ccw_edges_around_face, face, "lpmx_mesh_ccw_edges_around_face")

    def ccw_adjacent_faces(self, face):
        return self._index_list(_lib.lib().lpmx_mesh_ccw_adjacent_faces, face, "lpmx_mesh_ccw_adjacent_faces")

    def neighbors_flag(self, flags, start=0, end=None):
        """NeighborsFlag over faces [start, end): flags |= a neighbour is more than one level finer.  In place; returns the
        number of newly set flags."""
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        if f is not flags:
            raise ValueError("flags must be a C-contiguous uint8 array (updated in place)")
        n = ctypes.c_int()
        rc = _lib.lib().lpmx_mesh_neighbors_flag(self._m, f.ctypes.data, int(start), int(self.n_faces if end is None else end),
                                                 ctypes.byref(n))
        if rc:
            raise LpmxError(rc, "lpmx_mesh_neighbors_flag")
        return n.value

    def _locate(self, mode, pts, start=None):
        self._push_coordinates()
        p = np.ascontiguousarray(np.atleast_2d(np.asarray(pts, dtype=np.float64)))
        if p.shape[1] != self.ndim:
            raise ValueError(f"points must have {self.ndim} columns")
        out = np.empty(p.shape[0], dtype=np.int32)
        st = None if start is None else np.ascontiguousarray(np.broadcast_to(np.asarray(start, dtype=np.int32), (p.shape[0],)))
        rc = _lib.lib().lpmx_mesh_locate(self._m, mode, p.ctypes.data, p.shape[0], None if st is None else st.ctypes.data,
                                         out.ctypes.data)
        if rc:
            raise LpmxError(rc, "lpmx_mesh_locate")
        return out

    def locate_face_containing_pt(self, pts):
        return self._locate(0, pts)

    def locate_pt_walk_search(self, pts, face_start_idx):
        return self._locate(1, pts, face_start_idx)

    def locate_pt_tree_search(self, pts, root_face):
        return self._locate(2, pts, root_face)

    def nearest_root_face(self, pts):
        return self._locate(3, pts)

    # names used by the reference's examples
    def n_vertices_host(self):
        return self.n_verts

    def n_faces_host(self):
        return self.n_faces

    def appx_mesh_size(self):
        """Faces::appx_mesh_size (src/mesh/lpm_faces_impl.hpp:214-227)."""
        return float(np.sqrt(self.face_area[self.face_mask == 0].sum() / self.n_face_leaves))


def _ptr(a):
    """void* of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported array type {type(a)}")


def _f64(a):
    if a is None or hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    if a is None or hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.uint8)


def _empty_like_vec(ref, n, layout, ld):
    if hasattr(ref, "data_ptr"):
        import torch
        shape = (n, 3) if layout == LAYOUT_RIGHT else (3, ld)
        return torch.empty(shape, dtype=torch.float64, device=ref.device)
    return np.empty((n, 3) if layout == LAYOUT_RIGHT else (3, ld), dtype=np.float64)


def _empty_like_scalar(ref, n):
    if hasattr(ref, "data_ptr"):
        import torch
        return torch.empty(n, dtype=torch.float64, device=ref.device)
    return np.empty(n, dtype=np.float64)


def gmls_params(order_or_params=3, **overrides):
    """gmls::Params(order) (src/lpm_compadre.hpp:49-60) as a GmlsParams struct; keyword overrides set members."""
    if isinstance(order_or_params, _lib.GmlsParams):
        p = order_or_params
    else:
        p = _lib.GmlsParams()
        rc = _lib.lib().lpmx_gmls_params_init(ctypes.byref(p), int(order_or_params))
        if rc:
            raise LpmxError(rc, "lpmx_gmls_params_init")
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


class GmlsLaplacian:
    """The built-in surface-Laplacian provider (lpmx_gmls_swe_laplacian + lpmx_gmls_provider_t): pass an instance as
    `laplacian=` to SWESolver.advance / swe_rk2_step; the reference's gather -> Compadre -> scatter stays on the device."""

    def __init__(self, engine, params=None):
        self.provider = _lib.GmlsProvider()
        self.provider.handle = engine._h
        self.provider.params = gmls_params(params if params is not None else 3)
        self.fn = ctypes.cast(engine._L.lpmx_gmls_swe_laplacian, _lib.SWE_LAPLACIAN_FN)
        self.user = ctypes.cast(ctypes.pointer(self.provider), ctypes.c_void_p)


class Engine:
    """One engine handle per process and GPU (lpmx_create)."""

    def __init__(self, device_id=0):
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        rc = self._L.lpmx_create(ctypes.byref(h), device_id)
        if rc:
            raise LpmxError(rc, "lpmx_create", "(no CUDA device: this engine has no CPU fallback)")
        self._h = h
        self.device_id = device_id

    def close(self):
        if getattr(self, "_h", None):
            self._L.lpmx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc:
            raise LpmxError(rc, where, self._L.lpmx_last_error_string(self._h).decode())

    def sync(self):
        self._check(self._L.lpmx_sync(self._h), "lpmx_sync")

    def stream(self):
        s = ctypes.c_void_p()
        self._check(self._L.lpmx_stream(self._h, ctypes.byref(s)), "lpmx_stream")
        return s.value

    def launch_count(self):
        n = ctypes.c_long()
        self._check(self._L.lpmx_launch_count(self._h, ctypes.byref(n)), "lpmx_launch_count")
        return n.value

    def fp64_peak_tflops(self):
        t, ms = ctypes.c_double(), ctypes.c_double()
        self._check(self._L.lpmx_fp64_peak_tflops(self._h, ctypes.byref(t), ctypes.byref(ms)), "lpmx_fp64_peak_tflops")
        return t.value

    def profile_enable(self, enable=True):
        self._check(self._L.lpmx_profile_enable(self._h, int(bool(enable))), "lpmx_profile_enable")

    def profile_read(self):
        """(pair-sum launches, their summed device ms, pair visits incl. zero padding) since the last read."""
        n, ms, pv = ctypes.c_long(), ctypes.c_double(), ctypes.c_double()
        self._check(self._L.lpmx_profile_read(self._h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(pv)),
                    "lpmx_profile_read")
        return n.value, ms.value, pv.value

    def set_partition(self, rank, world):
        self._check(self._L.lpmx_set_partition(self._h, rank, world), "lpmx_set_partition")

    @staticmethod
    def comm_unique_id():
        buf = ctypes.create_string_buffer(128)
        rc = _lib.lib().lpmx_comm_unique_id(buf)
        if rc:
            raise LpmxError(rc, "lpmx_comm_unique_id")
        return bytes(buf.raw)

    def comm_init(self, unique_id, rank, world):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._L.lpmx_comm_init(self._h, buf, rank, world), "lpmx_comm_init")

    def const_stream_launch_count(self):
        """Bank-kernel launches issued so far (0: every pair sum went through the default kernel)."""
        n = ctypes.c_long(0)
        self._check(self._L.lpmx_const_stream_launch_count(self._h, ctypes.byref(n)), "lpmx_const_stream_launch_count")
        return int(n.value)

    def pair_sum_const_stream(self, mode):
        """0 off, 1 overlapped, 2 serial, -1 environment: velocity pair sums through the constant bank (include/lpmx.h)."""
        self._check(self._L.lpmx_pair_sum_const_stream(self._h, int(mode)), "lpmx_pair_sum_const_stream")

    def comm_enable_peer_exchange(self, enable=True):
        """Collective, after comm_init and before any solver exists: exchange the packed source records with one
        kernel that stores into the peers' slabs over NVLink instead of the NCCL broadcasts (include/lpmx.h).
        Raises LpmxError(LPMX_ERR_UNSUPPORTED) when the GPUs cannot map each other's memory."""
        self._check(self._L.lpmx_comm_enable_peer_exchange(self._h, 1 if enable else 0), "lpmx_comm_enable_peer_exchange")

    def comm_peer_exchange_enabled(self):
        """(enabled, number of solver slabs currently mapped into the peers)."""
        en, nr = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self._L.lpmx_comm_peer_exchange_enabled(self._h, ctypes.byref(en), ctypes.byref(nr)),
                    "lpmx_comm_peer_exchange_enabled")
        return bool(en.value), nr.value

    # ---- operator level -----------------------------------------------------------------
    @staticmethod
    def _vec_args(x, layout, ld):
        """(array, n, ld) for a Real*[3] view."""
        if x is None:
            return None, 0, 0
        if layout == LAYOUT_RIGHT:
            return x, int(x.shape[0]), 0
        return x, None, int(ld if ld else x.shape[1])

    def _sums_common(self, tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld):
        src_xyz = _f64(src_xyz)
        tgt_xyz = _f64(tgt_xyz)
        if layout == LAYOUT_RIGHT:
            n_src = src_xyz.shape[0] if n_src is None else n_src
            if tgt_xyz is not None:
                n_tgt = tgt_xyz.shape[0] if n_tgt is None else n_tgt
        else:
            src_ld = src_ld or src_xyz.shape[1]
            n_src = src_xyz.shape[1] if n_src is None else n_src
            if tgt_xyz is not None:
                tgt_ld = tgt_ld or tgt_xyz.shape[1]
                n_tgt = tgt_xyz.shape[1] if n_tgt is None else n_tgt
        if tgt_xyz is None:
            n_tgt, tgt_ld = n_src, src_ld
        return tgt_xyz, src_xyz, int(n_tgt), int(n_src), int(tgt_ld or 0), int(src_ld or 0)

    def bve_velocity(self, tgt_xyz, src_xyz, src_vort, src_area, src_mask, collocated=False, layout=LAYOUT_RIGHT,
                     out=None, n_tgt=None, n_src=None, tgt_ld=0, src_ld=0):
        """BVEVertexVelocity (collocated=False) / BVEFaceVelocity (collocated=True, tgt_xyz=None)."""
        tgt_xyz, src_xyz, n_tgt, n_src, tgt_ld, src_ld = self._sums_common(
            None if collocated else tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld)
        src_vort, src_area, src_mask = _f64(src_vort), _f64(src_area), _u8(src_mask)
        if out is None:
            out = _empty_like_vec(src_xyz, n_tgt, layout, tgt_ld)
        self._check(self._L.lpmx_bve_velocity(self._h, _ptr(tgt_xyz), layout, tgt_ld, n_tgt, _ptr(src_xyz), layout,
                                              src_ld, _ptr(src_vort), _ptr(src_area), _ptr(src_mask), n_src,
                                              int(bool(collocated)), _ptr(out)), "lpmx_bve_velocity")
        return out

    def bve_streamfn(self, tgt_xyz, src_xyz, src_vort, src_area, src_mask, collocated=False, layout=LAYOUT_RIGHT,
                     out=None, n_tgt=None, n_src=None, tgt_ld=0, src_ld=0):
        """BVEVertexStreamFn / BVEFaceStreamFn."""
        tgt_xyz, src_xyz, n_tgt, n_src, tgt_ld, src_ld = self._sums_common(
            None if collocated else tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld)
        src_vort, src_area, src_mask = _f64(src_vort), _f64(src_area), _u8(src_mask)
        if out is None:
            out = _empty_like_scalar(src_xyz, n_tgt)
        self._check(self._L.lpmx_bve_streamfn(self._h, _ptr(tgt_xyz), layout, tgt_ld, n_tgt, _ptr(src_xyz), layout,
                                              src_ld, _ptr(src_vort), _ptr(src_area), _ptr(src_mask), n_src,
                                              int(bool(collocated)), _ptr(out)), "lpmx_bve_streamfn")
        return out

    def set_io_sharded(self, on=True):
        """Host arrays of the BVE / IC2D solvers and in-place steppers carry only this rank's rows (lpmx_set_io_sharded)."""
        self._check(self._L.lpmx_set_io_sharded(self._h, int(bool(on))), "lpmx_set_io_sharded")

    def local_targets(self, n_first, n_second, mask_second):
        """(A, B): this rank's leaf faces and its non-source targets, as indices into [vertices | faces] (lpmx_local_targets)."""
        mask = np.ascontiguousarray(mask_second, dtype=np.uint8)
        idx = np.zeros(n_first + n_second, dtype=np.int32)
        na, nb = ctypes.c_int(), ctypes.c_int()
        self._check(self._L.lpmx_local_targets(self._h, n_first, n_second, _ptr(mask), _ptr(idx), ctypes.byref(na),
                                               ctypes.byref(nb)), "lpmx_local_targets")
        return idx[:na.value].copy(), idx[na.value:na.value + nb.value].copy()

    def bve_solve(self, tgt_xyz, src_xyz, src_vort, src_area, src_mask, collocated=False, layout=LAYOUT_RIGHT,
                  n_tgt=None, n_src=None, tgt_ld=0, src_ld=0):
        """BVEVertexSolve / BVEFaceSolve: (psi, velocity) of the targets in one pass of the fused pair kernel."""
        tgt_xyz, src_xyz, n_tgt, n_src, tgt_ld, src_ld = self._sums_common(
            None if collocated else tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld)
        src_vort, src_area, src_mask = _f64(src_vort), _f64(src_area), _u8(src_mask)
        vel = _empty_like_vec(src_xyz, n_tgt, layout, tgt_ld)
        psi = _empty_like_scalar(src_xyz, n_tgt)
        self._check(self._L.lpmx_bve_solve(self._h, _ptr(tgt_xyz), layout, tgt_ld, n_tgt, _ptr(src_xyz), layout, src_ld,
                                           _ptr(src_vort), _ptr(src_area), _ptr(src_mask), n_src, int(bool(collocated)),
                                           _ptr(psi), _ptr(vel)), "lpmx_bve_solve")
        return psi, vel

    def ic2d_sums(self, tgt_xyz, src_xyz, src_vort, src_area, src_mask, eps=0.0, targets_are_sources=False,
                  with_psi=True, layout=LAYOUT_RIGHT, n_tgt=None, n_src=None, tgt_ld=0, src_ld=0):
        """Incompressible2DPassiveSums (targets_are_sources=False) / ActiveSums (True)."""
        tgt_xyz, src_xyz, n_tgt, n_src, tgt_ld, src_ld = self._sums_common(
            None if targets_are_sources else tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld)
        src_vort, src_area, src_mask = _f64(src_vort), _f64(src_area), _u8(src_mask)
        vel = _empty_like_vec(src_xyz, n_tgt, layout, tgt_ld)
        psi = _empty_like_scalar(src_xyz, n_tgt) if with_psi else None
        self._check(self._L.lpmx_ic2d_sums(self._h, _ptr(tgt_xyz), layout, tgt_ld, n_tgt, _ptr(src_xyz), layout, src_ld,
                                           _ptr(src_vort), _ptr(src_area), _ptr(src_mask), n_src, float(eps),
                                           int(bool(targets_are_sources)), _ptr(vel), _ptr(psi)), "lpmx_ic2d_sums")
        return vel, psi

    def swe_sphere_sums(self, tgt_xyz, src_xyz, src_vort, src_div, src_area, src_mask, eps=0.0,
                        targets_are_sources=False, do_velocity=True, want_grad=False, layout=LAYOUT_RIGHT,
                        n_tgt=None, n_src=None, tgt_ld=0, src_ld=0):
        """SphereVertexSums (targets_are_sources=False) / SphereFaceSums (True)."""
        tgt_xyz, src_xyz, n_tgt, n_src, tgt_ld, src_ld = self._sums_common(
            None if targets_are_sources else tgt_xyz, src_xyz, layout, n_tgt, n_src, tgt_ld, src_ld)
        src_vort, src_div, src_area, src_mask = _f64(src_vort), _f64(src_div), _f64(src_area), _u8(src_mask)
        vel = _empty_like_vec(src_xyz, n_tgt, layout, tgt_ld) if do_velocity else None
        ddot = _empty_like_scalar(src_xyz, n_tgt)
        grad = None
        if want_grad:
            grad = _empty_like_scalar(src_xyz, 9 * n_tgt)
        self._check(self._L.lpmx_swe_sphere_sums(self._h, _ptr(tgt_xyz), layout, tgt_ld, n_tgt, _ptr(src_xyz), layout,
                                                 src_ld, _ptr(src_vort), _ptr(src_div), _ptr(src_area), _ptr(src_mask),
                                                 n_src, float(eps), int(bool(targets_are_sources)),
                                                 int(bool(do_velocity)), _ptr(vel), _ptr(ddot), _ptr(grad)),
                    "lpmx_swe_sphere_sums")
        if want_grad:
            grad = grad.reshape(n_tgt, 9)
        return vel, ddot, grad

    # ---- diagnostics ---------------------------------------------------------------------
    def ic2d_totals(self, active_vort, active_vel, active_area, active_mask, layout=LAYOUT_RIGHT):
        """(total_vorticity, total_kinetic_energy, total_enstrophy) -- Incompressible2D::total_*."""
        n = active_vel.shape[0] if layout == LAYOUT_RIGHT else active_vel.shape[1]
        v, k, e = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._check(self._L.lpmx_ic2d_totals(self._h, n, _ptr(_f64(active_vort)), _ptr(_f64(active_vel)), layout, n,
                                             _ptr(_f64(active_area)), _ptr(_u8(active_mask)), ctypes.byref(v),
                                             ctypes.byref(k), ctypes.byref(e)), "lpmx_ic2d_totals")
        return v.value, k.value, e.value

    def err_norms(self, err, exact, weight, layout=LAYOUT_RIGHT):
        """ErrNorms(err, exact, weight) -> (l1, l2, linf)."""
        err, exact, weight = _f64(err), _f64(exact), _f64(weight)
        ndim = 1 if err.ndim == 1 else 3
        n = err.shape[0] if (ndim == 1 or layout == LAYOUT_RIGHT) else err.shape[1]
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._check(self._L.lpmx_err_norms(self._h, n, ndim, _ptr(err), _ptr(exact), layout, n, _ptr(weight),
                                           ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "lpmx_err_norms")
        return a.value, b.value, c.value

    def ftle(self, geom, vert_phys, vert_ref, face_phys, face_ref, face_verts, face_mask, ftle=None,
             layout=LAYOUT_RIGHT, verts_layout=LAYOUT_RIGHT):
        """ComputeFTLE<Seed> + get_max_ftle (src/mesh/lpm_ftle.hpp) for quadrilateral faces; geom 0 = sphere, 1 = plane.
        `face_phys` is normalised in place on the sphere, `ftle` (zeros if omitted) is written at the leaves only.
        Returns (ftle, max_ftle)."""
        vert_phys, vert_ref, face_ref = _f64(vert_phys), _f64(vert_ref), _f64(face_ref)
        if isinstance(face_phys, np.ndarray) and (face_phys.dtype != np.float64 or not face_phys.flags.c_contiguous):
            raise ValueError("face_phys is an in/out argument: pass a C-contiguous float64 array")
        nd = 3 if geom == 0 else 2
        nv = vert_ref.shape[0] if layout == LAYOUT_RIGHT else vert_ref.shape[1]
        nf = face_ref.shape[0] if layout == LAYOUT_RIGHT else face_ref.shape[1]
        assert (vert_ref.shape[1] if layout == LAYOUT_RIGHT else vert_ref.shape[0]) == nd
        fv = face_verts if hasattr(face_verts, "data_ptr") else np.ascontiguousarray(face_verts, dtype=np.int32)
        if ftle is None:
            ftle = _empty_like_scalar(face_ref, nf)
            ftle[...] = 0.0
        mx = ctypes.c_double()
        self._check(self._L.lpmx_ftle(self._h, geom, nv, _ptr(vert_phys), _ptr(vert_ref), layout, nv, nf, _ptr(face_phys),
                                      _ptr(face_ref), nf, _ptr(fv), verts_layout, _ptr(_u8(face_mask)), _ptr(ftle),
                                      ctypes.byref(mx)), "lpmx_ftle")
        return ftle, mx.value

    # ---- adaptive-refinement flags (src/mesh/lpm_refinement_flags.hpp, lpm_refinement.hpp) ----
    FLAG_KINDS = {"scalar_max": 0, "scalar_integral": 1, "scalar_variation": 2, "flow_map_variation": 3}

    def _flag_desc(self, kind, face_mask, face_vals=None, area=None, vert_vals=None, face_verts=None, vert_lag=None, tol=0.0):
        keep = []  # arrays must outlive the call
        d = _lib.FlagDesc()
        d.kind = self.FLAG_KINDS[kind] if isinstance(kind, str) else int(kind)
        mask = _u8(face_mask)
        keep.append(mask)
        d.n_faces = mask.shape[0]
        d.mask = _ptr(mask).value if d.n_faces else None

        def put(field, a):
            if a is not None:
                keep.append(a)
                setattr(d, field, _ptr(a).value if a.shape[0] else None)
        put("face_vals", _f64(face_vals))
        put("area", _f64(area))
        vv = _f64(vert_vals)
        put("vert_vals", vv)
        if face_verts is not None:
            fv = face_verts if hasattr(face_verts, "data_ptr") else np.ascontiguousarray(face_verts, dtype=np.int32)
            d.n_face_verts = fv.shape[1]
            put("face_verts", fv)
        vl = _f64(vert_lag)
        if vl is not None:
            d.ndim, d.layout, d.ld = vl.shape[1], LAYOUT_RIGHT, vl.shape[0]
            put("vert_lag", vl)
        d.n_verts = vv.shape[0] if vv is not None else (vl.shape[0] if vl is not None else 0)
        d.tol = float(tol)
        return d, keep

    def refine_flag_max(self, kind, face_mask, **arrays):
        """The reduction of <Flag>::set_tol_from_relative_value(): tol = relative_tol * refine_flag_max(...)."""
        d, keep = self._flag_desc(kind, face_mask, **arrays)
        mx = ctypes.c_double()
        self._check(self._L.lpmx_refine_flag_max(self._h, ctypes.byref(d), ctypes.byref(mx)), "lpmx_refine_flag_max")
        return mx.value

    def refine_flag(self, kind, face_mask, tol, start=0, end=None, flags=None, **arrays):
        """Refinement::iterate with one flag functor over faces [start, end): returns (flags, count).  With flags=None
        the flags start cleared; an existing uint8 array is updated in place (flags are only ever switched on)."""
        d, keep = self._flag_desc(kind, face_mask, tol=tol, **arrays)
        end = d.n_faces if end is None else end
        clear = flags is None
        if flags is None:
            flags = np.zeros(d.n_faces, dtype=np.uint8)
        elif isinstance(flags, np.ndarray) and (flags.dtype != np.uint8 or not flags.flags.c_contiguous):
            raise ValueError("flags is an in/out argument: pass a C-contiguous uint8 array")
        ct = ctypes.c_int()
        self._check(self._L.lpmx_refine_flag(self._h, ctypes.byref(d), start, end, 1 if clear else 0, _ptr(flags),
                                             ctypes.byref(ct)), "lpmx_refine_flag")
        return flags, ct.value

    # ---- gather / scatter and the GMLS surface Laplacian (the steps either side of the SWE sums) ----
    def gather_mesh_data(self, vert_data, face_data, face_mask):
        """GatherMeshData: vertices then leaf faces (row n_verts + leaf_idx(f)); 1-D or (n, k) arrays, LayoutRight."""
        vert_data, face_data = _f64(vert_data), _f64(face_data)
        ncomp = 1 if face_data.ndim == 1 else face_data.shape[1]
        nv, nf = vert_data.shape[0], face_data.shape[0]
        mask = _u8(face_mask)
        n = ctypes.c_int()
        self._check(self._L.lpmx_gather_mesh_data(self._h, ncomp, LAYOUT_RIGHT, nv, _ptr(vert_data), nv, nf, _ptr(face_data), nf,
                                                  _ptr(mask), None, 0, ctypes.byref(n)), "lpmx_gather_mesh_data")
        out = np.empty((n.value,) if face_data.ndim == 1 else (n.value, ncomp))
        self._check(self._L.lpmx_gather_mesh_data(self._h, ncomp, LAYOUT_RIGHT, nv, _ptr(vert_data), nv, nf, _ptr(face_data), nf,
                                                  _ptr(mask), _ptr(out), n.value, ctypes.byref(n)), "lpmx_gather_mesh_data")
        return out

    def scatter_mesh_data(self, gathered, vert_out, face_out, face_mask):
        """ScatterMeshData::scatter_fields into vert_out / face_out (float64, C-contiguous; divided faces untouched)."""
        gathered = _f64(gathered)
        ncomp = 1 if gathered.ndim == 1 else gathered.shape[1]
        nv, nf = vert_out.shape[0], face_out.shape[0]
        self._check(self._L.lpmx_scatter_mesh_data(self._h, ncomp, LAYOUT_RIGHT, _ptr(gathered), gathered.shape[0], nv,
                                                   _ptr(vert_out), nv, nf, _ptr(face_out), nf, _ptr(_u8(face_mask))),
                    "lpmx_scatter_mesh_data")

    def gmls_sphere_laplacian(self, xyz, f, params=None, layout=LAYOUT_RIGHT, diagnostics=False):
        """Surface Laplacian of the samples f at the points xyz (lpmx_gmls_sphere_laplacian); params: GmlsParams or an
        int order (gmls::Params(order)).  Returns lap, or (lap, window_radius, n_neighbors) with diagnostics=True."""
        params = gmls_params(params if params is not None else 3)
        xyz, f = _f64(xyz), _f64(f)
        n = f.shape[0]
        lap = _empty_like_scalar(f, n)
        eps = _empty_like_scalar(f, n) if diagnostics else None
        if diagnostics and hasattr(f, "data_ptr"):
            import torch
            nn = torch.empty(n, dtype=torch.int32, device=f.device)
        else:
            nn = np.empty(n, dtype=np.int32) if diagnostics else None
        self._check(self._L.lpmx_gmls_sphere_laplacian(self._h, ctypes.byref(params), n, _ptr(xyz), layout, n, _ptr(f), _ptr(lap),
                                                       _ptr(eps), _ptr(nn)), "lpmx_gmls_sphere_laplacian")
        return (lap, eps, nn) if diagnostics else lap

    def gmls_sphere_interpolate(self, src_xyz, src_fields, tgt_xyz, params=None):
        """Scalar GMLS point evaluation (remeshing): src_fields is a list of n_src arrays (or an (n_fields, n_src)
        array); returns an (n_fields, n_tgt) array -- lpmx_gmls_sphere_interpolate.  numpy in, numpy out."""
        params = gmls_params(params if params is not None else 3)
        src_xyz, tgt_xyz = _f64(src_xyz), _f64(tgt_xyz)
        F = [np.ascontiguousarray(f, dtype=np.float64) for f in src_fields]
        nf, ns, nt = len(F), src_xyz.shape[0], tgt_xyz.shape[0]
        out = np.empty((nf, nt))
        pin = (ctypes.c_void_p * nf)(*[f.ctypes.data for f in F])
        pout = (ctypes.c_void_p * nf)(*[out[k].ctypes.data for k in range(nf)])
        self._check(self._L.lpmx_gmls_sphere_interpolate(self._h, ctypes.byref(params), ns, _ptr(src_xyz), LAYOUT_RIGHT, ns, nf,
                                                         pin, nt, _ptr(tgt_xyz), LAYOUT_RIGHT, nt, pout),
                    "lpmx_gmls_sphere_interpolate")
        return out

    def gmls_provider(self, params=None):
        """(callback, user) for SWESolver.advance / swe_rk2_step: the built-in device-side surface-Laplacian provider."""
        return GmlsLaplacian(self, params)

    # ---- stepper level, in place on caller arrays ----------------------------------------
    def bve_rk4_step(self, dt, Omega, vert_xyz, vert_vort, vert_vel, face_xyz, face_vort, face_vel, face_area,
                     face_mask, n_steps=1, layout=LAYOUT_RIGHT):
        """BVERK4::advance_timestep, n_steps times, in place (arrays must be C-contiguous float64)."""
        nv = vert_xyz.shape[0] if layout == LAYOUT_RIGHT else vert_xyz.shape[1]
        nf = face_xyz.shape[0] if layout == LAYOUT_RIGHT else face_xyz.shape[1]
        self._check(self._L.lpmx_bve_rk4_step(self._h, float(dt), float(Omega), nv, _ptr(vert_xyz), _ptr(vert_vort),
                                              _ptr(vert_vel), nf, _ptr(face_xyz), _ptr(face_vort), _ptr(face_vel),
                                              _ptr(face_area), _ptr(face_mask), layout, nv, nf, n_steps),
                    "lpmx_bve_rk4_step")

    def ic2d_rk2_step(self, dt, Omega, eps, passive_xyz, passive_vort, passive_vel, passive_psi, active_xyz,
                      active_vort, active_vel, active_psi, active_area, active_mask, n_steps=1, layout=LAYOUT_RIGHT):
        """Incompressible2DRK2::advance_timestep_impl, n_steps times, in place."""
        np_ = passive_xyz.shape[0] if layout == LAYOUT_RIGHT else passive_xyz.shape[1]
        na = active_xyz.shape[0] if layout == LAYOUT_RIGHT else active_xyz.shape[1]
        self._check(self._L.lpmx_ic2d_rk2_step(self._h, float(dt), float(Omega), float(eps), np_, _ptr(passive_xyz),
                                               _ptr(passive_vort), _ptr(passive_vel), _ptr(passive_psi), na,
                                               _ptr(active_xyz), _ptr(active_vort), _ptr(active_vel), _ptr(active_psi),
                                               _ptr(active_area), _ptr(active_mask), layout, np_, na, n_steps),
                    "lpmx_ic2d_rk2_step")


class BVESolver:
    """Device-resident BVESphere state + BVERK4 (lpmx_bve_solver_*)."""

    def __init__(self, engine, n_verts, n_faces):
        self.e = engine
        self.nv, self.nf = n_verts, n_faces
        s = ctypes.c_void_p()
        engine._check(engine._L.lpmx_bve_solver_create(engine._h, n_verts, n_faces, ctypes.byref(s)),
                      "lpmx_bve_solver_create")
        self._s = s

    def close(self):
        if getattr(self, "_s", None) and getattr(self.e, "_h", None):
            self.e._L.lpmx_bve_solver_destroy(self._s)
        self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, vert_xyz, vert_vort, vert_vel, face_xyz, face_vort, face_vel, face_area, face_mask,
                  layout=LAYOUT_RIGHT):
        self.e._check(self.e._L.lpmx_bve_solver_set_state(
            self._s, _ptr(vert_xyz), _ptr(vert_vort), _ptr(vert_vel), _ptr(face_xyz), _ptr(face_vort), _ptr(face_vel),
            _ptr(face_area), _ptr(face_mask), layout, self.nv, self.nf), "lpmx_bve_solver_set_state")

    def get_state(self, vert_xyz=None, vert_vort=None, vert_vel=None, face_xyz=None, face_vort=None, face_vel=None,
                  layout=LAYOUT_RIGHT):
        self.e._check(self.e._L.lpmx_bve_solver_get_state(
            self._s, _ptr(vert_xyz), _ptr(vert_vort), _ptr(vert_vel), _ptr(face_xyz), _ptr(face_vort), _ptr(face_vel),
            layout, self.nv, self.nf), "lpmx_bve_solver_get_state")

    def init_velocity(self):
        self.e._check(self.e._L.lpmx_bve_solver_init_velocity(self._s), "lpmx_bve_solver_init_velocity")

    def stream_fn(self, vert_psi, face_psi):
        self.e._check(self.e._L.lpmx_bve_solver_stream_fn(self._s, _ptr(vert_psi), _ptr(face_psi)),
                      "lpmx_bve_solver_stream_fn")

    def advance(self, dt, Omega, n_steps=1):
        self.e._check(self.e._L.lpmx_bve_solver_advance(self._s, float(dt), float(Omega), n_steps),
                      "lpmx_bve_solver_advance")

    def interactions_per_eval(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self.e._check(self.e._L.lpmx_bve_solver_interactions_per_eval(self._s, ctypes.byref(a), ctypes.byref(b)),
                      "lpmx_bve_solver_interactions_per_eval")
        return a.value, b.value


class IC2DSolver:
    """Device-resident Incompressible2D state + Incompressible2DRK2 (lpmx_ic2d_solver_*)."""

    def __init__(self, engine, n_passive, n_active, eps=0.0):
        self.e = engine
        self.np_, self.na = n_passive, n_active
        s = ctypes.c_void_p()
        engine._check(engine._L.lpmx_ic2d_solver_create(engine._h, n_passive, n_active, float(eps), ctypes.byref(s)),
                      "lpmx_ic2d_solver_create")
        self._s = s

    def close(self):
        if getattr(self, "_s", None) and getattr(self.e, "_h", None):
            self.e._L.lpmx_ic2d_solver_destroy(self._s)
        self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, passive_xyz, passive_vort, passive_vel, active_xyz, active_vort, active_vel, active_area,
                  active_mask, layout=LAYOUT_RIGHT):
        self.e._check(self.e._L.lpmx_ic2d_solver_set_state(
            self._s, _ptr(passive_xyz), _ptr(passive_vort), _ptr(passive_vel), _ptr(active_xyz), _ptr(active_vort),
            _ptr(active_vel), _ptr(active_area), _ptr(active_mask), layout, self.np_, self.na),
            "lpmx_ic2d_solver_set_state")

    def get_state(self, passive_xyz=None, passive_vort=None, passive_vel=None, passive_psi=None, active_xyz=None,
                  active_vort=None, active_vel=None, active_psi=None, layout=LAYOUT_RIGHT):
        self.e._check(self.e._L.lpmx_ic2d_solver_get_state(
            self._s, _ptr(passive_xyz), _ptr(passive_vort), _ptr(passive_vel), _ptr(passive_psi), _ptr(active_xyz),
            _ptr(active_vort), _ptr(active_vel), _ptr(active_psi), layout, self.np_, self.na),
            "lpmx_ic2d_solver_get_state")

    def init_direct_sums(self):
        self.e._check(self.e._L.lpmx_ic2d_solver_init_direct_sums(self._s), "lpmx_ic2d_solver_init_direct_sums")

    def totals(self):
        """(total_vorticity, total_kinetic_energy, total_enstrophy) of the resident state."""
        v, k, e = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self.e._check(self.e._L.lpmx_ic2d_solver_totals(self._s, ctypes.byref(v), ctypes.byref(k), ctypes.byref(e)),
                      "lpmx_ic2d_solver_totals")
        return v.value, k.value, e.value

    def lazy_stream_fn(self, demand_next):
        """Override the lazy-psi heuristic for the next advance (True: fuse psi into its final evaluation)."""
        self.e._check(self.e._L.lpmx_ic2d_solver_lazy_stream_fn(self._s, int(bool(demand_next))), "lpmx_ic2d_solver_lazy_stream_fn")

    def advance(self, dt, Omega, n_steps=1):
        self.e._check(self.e._L.lpmx_ic2d_solver_advance(self._s, float(dt), float(Omega), n_steps),
                      "lpmx_ic2d_solver_advance")


PASSIVE_FIELDS = ("xyz", "vort", "div", "depth", "surf", "bottom", "vel", "ddot", "laps")
ACTIVE_FIELDS = ("xyz", "vort", "div", "area", "mass", "depth", "surf", "bottom", "vel", "ddot", "laps")


def _swe_structs(passive, active, mask):
    """dicts of arrays (missing / None = NULL) -> (lpmx_swe_passive_t, lpmx_swe_active_t, keep-alive list)."""
    P, A = _lib.SwePassive(), _lib.SweActive()
    for k in PASSIVE_FIELDS:
        a = passive.get(k)
        setattr(P, k, None if a is None else _ptr(a))
    for k in ACTIVE_FIELDS:
        a = active.get(k)
        setattr(A, k, None if a is None else _ptr(a))
    A.mask = None if mask is None else _ptr(mask)
    return P, A


def _laplacian_cb(fn):
    """Wrap a Python callable fn(stage, stream, n_passive, passive_xyz_ptr, passive_surf_ptr, passive_laps_ptr,
    n_active, active_xyz_ptr, active_surf_ptr, active_mask_ptr, active_laps_ptr, xyz_ld) -> None (device pointers
    as ints) into an lpmx_swe_laplacian_fn; None -> NULL provider."""
    if fn is None:
        return ctypes.cast(None, _lib.SWE_LAPLACIAN_FN)
    if isinstance(fn, GmlsLaplacian):
        return fn.fn

    def cb(user, stage, stream, n_p, pxyz, psurf, plaps, n_a, axyz, asurf, amask, alaps, ld):
        try:
            fn(stage, stream, n_p, pxyz, psurf, plaps, n_a, axyz, asurf, amask, alaps, ld)
            return 0
        except Exception:  # a Python exception must not unwind through C
            import traceback
            traceback.print_exc()
            return 1
    return _lib.SWE_LAPLACIAN_FN(cb)


def swe_rk2_step(engine, dt, Omega, g, eps, passive, active, mask, laplacian=None, n_steps=1, layout=LAYOUT_RIGHT):
    """SWERK2::advance_timestep_impl, n_steps times, in place on the arrays in the dicts `passive` / `active`
    (keys PASSIVE_FIELDS / ACTIVE_FIELDS; float64, C-contiguous) -- lpmx_swe_rk2_step."""
    nv = passive["xyz"].shape[0] if layout == LAYOUT_RIGHT else passive["xyz"].shape[1]
    nf = active["xyz"].shape[0] if layout == LAYOUT_RIGHT else active["xyz"].shape[1]
    P, A = _swe_structs(passive, active, mask)
    cb = _laplacian_cb(laplacian)
    engine._check(engine._L.lpmx_swe_rk2_step(engine._h, float(dt), float(Omega), float(g), float(eps), nv,
                                              ctypes.byref(P), nf, ctypes.byref(A), layout, nv, nf, cb,
                                              getattr(laplacian, "user", None), n_steps),
                  "lpmx_swe_rk2_step")


class SWESolver:
    """Device-resident SWE<Seed> fields + SWERK2 (lpmx_swe_solver_*)."""

    def __init__(self, engine, n_passive, n_active, eps=0.0):
        self.e = engine
        self.np_, self.na = n_passive, n_active
        s = ctypes.c_void_p()
        engine._check(engine._L.lpmx_swe_solver_create(engine._h, n_passive, n_active, float(eps), ctypes.byref(s)),
                      "lpmx_swe_solver_create")
        self._s = s

    def close(self):
        if getattr(self, "_s", None) and getattr(self.e, "_h", None):
            self.e._L.lpmx_swe_solver_destroy(self._s)
        self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, passive, active, mask, layout=LAYOUT_RIGHT):
        P, A = _swe_structs(passive, active, mask)
        self.e._check(self.e._L.lpmx_swe_solver_set_state(self._s, ctypes.byref(P), ctypes.byref(A), layout, self.np_,
                                                          self.na), "lpmx_swe_solver_set_state")

    def get_state(self, passive, active, layout=LAYOUT_RIGHT):
        P, A = _swe_structs(passive, active, None)
        self.e._check(self.e._L.lpmx_swe_solver_get_state(self._s, ctypes.byref(P), ctypes.byref(A), layout, self.np_,
                                                          self.na), "lpmx_swe_solver_get_state")

    def set_laplacian(self, passive_laps, active_laps):
        self.e._check(self.e._L.lpmx_swe_solver_set_laplacian(self._s, _ptr(passive_laps), _ptr(active_laps)),
                      "lpmx_swe_solver_set_laplacian")

    def init_direct_sums(self, do_velocity=True):
        self.e._check(self.e._L.lpmx_swe_solver_init_direct_sums(self._s, int(bool(do_velocity))),
                      "lpmx_swe_solver_init_direct_sums")

    def advance(self, dt, Omega, g, laplacian=None, n_steps=1):
        cb = _laplacian_cb(laplacian)
        self.e._check(self.e._L.lpmx_swe_solver_advance(self._s, float(dt), float(Omega), float(g), cb,
                                                        getattr(laplacian, "user", None), n_steps),
                      "lpmx_swe_solver_advance")


# ---------------------------------------------------------------------------------------------------------
# Planar problems (PlaneGeometry): Real*[2] views
# ---------------------------------------------------------------------------------------------------------
TOPO_ZERO, TOPO_PLANAR_GAUSSIAN_MOUNTAIN = 0, 1
PLANE_PASSIVE_FIELDS, PLANE_ACTIVE_FIELDS = _lib.PLANE_PASSIVE_FIELDS, _lib.PLANE_ACTIVE_FIELDS
PLANE_SUM_FIELDS = _lib.PLANE_SUM_FIELDS


def _n2(x, layout):
    return int(x.shape[0] if layout == LAYOUT_RIGHT else x.shape[1])


def _empty_vec2(ref, n, layout):
    shape = (n, 2) if layout == LAYOUT_RIGHT else (2, n)
    if hasattr(ref, "data_ptr"):
        import torch
        return torch.empty(shape, dtype=torch.float64, device=ref.device)
    return np.empty(shape, dtype=np.float64)


def ic2d_plane_sums(engine, tgt_xy, src_xy, src_vort, src_area, src_mask, eps=0.0, targets_are_sources=False,
                    with_psi=True, layout=LAYOUT_RIGHT):
    """Incompressible2DPassiveSums<PlaneGeometry> (targets_are_sources=False) / ActiveSums (True) -- lpmx_ic2d_plane_sums."""
    src_xy, src_vort, src_area, src_mask = _f64(src_xy), _f64(src_vort), _f64(src_area), _u8(src_mask)
    n_src = _n2(src_xy, layout)
    tgt_xy = None if targets_are_sources else _f64(tgt_xy)
    n_tgt = n_src if targets_are_sources else _n2(tgt_xy, layout)
    vel = _empty_vec2(src_xy, n_tgt, layout)
    psi = _empty_like_scalar(src_xy, n_tgt) if with_psi else None
    engine._check(engine._L.lpmx_ic2d_plane_sums(engine._h, _ptr(tgt_xy), layout, n_tgt, n_tgt, _ptr(src_xy), layout, n_src,
                                                 _ptr(src_vort), _ptr(src_area), _ptr(src_mask), n_src, float(eps),
                                                 int(bool(targets_are_sources)), _ptr(vel), _ptr(psi)),
                  "lpmx_ic2d_plane_sums")
    return vel, psi


def ic2d_plane_rk2_step(engine, dt, f0, beta, eps, passive_xy, passive_vort, passive_vel, passive_psi, active_xy,
                        active_vort, active_vel, active_psi, active_area, active_mask, n_steps=1, layout=LAYOUT_RIGHT):
    """Incompressible2DRK2::advance_timestep_impl in the plane, n_steps times, in place -- lpmx_ic2d_plane_rk2_step."""
    np_, na = _n2(passive_xy, layout), _n2(active_xy, layout)
    engine._check(engine._L.lpmx_ic2d_plane_rk2_step(engine._h, float(dt), float(f0), float(beta), float(eps), np_,
                                                     _ptr(passive_xy), _ptr(passive_vort), _ptr(passive_vel),
                                                     _ptr(passive_psi), na, _ptr(active_xy), _ptr(active_vort),
                                                     _ptr(active_vel), _ptr(active_psi), _ptr(active_area),
                                                     _ptr(active_mask), layout, np_, na, n_steps),
                  "lpmx_ic2d_plane_rk2_step")


def swe_plane_sums(engine, tgt_xy, tgt_surf, src_xy, src_vort, src_div, src_area, src_mask, src_surf, eps, pse_eps,
                   targets_are_sources=False, do_velocity=True, layout=LAYOUT_RIGHT):
    """PlanarSWEVertexSums / PlanarSWEFaceSums -- lpmx_swe_plane_sums.  Returns a dict keyed by PLANE_SUM_FIELDS."""
    src_xy, src_vort, src_div, src_area, src_surf = map(_f64, (src_xy, src_vort, src_div, src_area, src_surf))
    src_mask = _u8(src_mask)
    n_src = _n2(src_xy, layout)
    if targets_are_sources:
        tgt_xy, tgt_surf, n_tgt = None, None, n_src
    else:
        tgt_xy, tgt_surf = _f64(tgt_xy), _f64(tgt_surf)
        n_tgt = _n2(tgt_xy, layout)
    out = {k: _empty_like_scalar(src_xy, n_tgt) for k in PLANE_SUM_FIELDS if k != "vel"}
    out["vel"] = _empty_vec2(src_xy, n_tgt, layout) if do_velocity else None
    S = _lib.PlaneSweSums()
    for k in PLANE_SUM_FIELDS:
        setattr(S, k, None if out[k] is None else _ptr(out[k]))
    engine._check(engine._L.lpmx_swe_plane_sums(engine._h, _ptr(tgt_xy), layout, n_tgt, _ptr(tgt_surf), n_tgt, _ptr(src_xy),
                                                layout, n_src, _ptr(src_vort), _ptr(src_div), _ptr(src_area),
                                                _ptr(src_mask), _ptr(src_surf), n_src, float(eps), float(pse_eps),
                                                int(bool(targets_are_sources)), int(bool(do_velocity)), ctypes.byref(S)),
                  "lpmx_swe_plane_sums")
    return out


def _plane_swe_structs(passive, active, mask):
    P, A = _lib.PlaneSwePassive(), _lib.PlaneSweActive()
    for k in PLANE_PASSIVE_FIELDS:
        a = passive.get(k)
        setattr(P, k, None if a is None else _ptr(a))
    for k in PLANE_ACTIVE_FIELDS:
        a = active.get(k)
        setattr(A, k, None if a is None else _ptr(a))
    A.mask = None if mask is None else _ptr(mask)
    return P, A


def swe_plane_rk4_step(engine, dt, f0, beta, g, eps, pse_eps, topo, passive, active, mask, n_steps=1,
                       layout=LAYOUT_RIGHT):
    """SWERK4::advance_timestep in the plane, n_steps times, in place on the arrays in the dicts (keys
    PLANE_PASSIVE_FIELDS / PLANE_ACTIVE_FIELDS; float64, C-contiguous) -- lpmx_swe_plane_rk4_step."""
    nv, nf = _n2(passive["xy"], layout), _n2(active["xy"], layout)
    P, A = _plane_swe_structs(passive, active, mask)
    engine._check(engine._L.lpmx_swe_plane_rk4_step(engine._h, float(dt), float(f0), float(beta), float(g), float(eps),
                                                    float(pse_eps), int(topo), nv, ctypes.byref(P), nf, ctypes.byref(A),
                                                    layout, nv, nf, n_steps), "lpmx_swe_plane_rk4_step")


class PlaneSWESolver:
    """Device-resident planar SWE<Seed> fields + SWERK4 (lpmx_plane_swe_solver_*)."""

    def __init__(self, engine, n_passive, n_active, eps, pse_eps, topo=TOPO_ZERO):
        self.e = engine
        self.np_, self.na = n_passive, n_active
        s = ctypes.c_void_p()
        engine._check(engine._L.lpmx_plane_swe_solver_create(engine._h, n_passive, n_active, float(eps), float(pse_eps),
                                                             int(topo), ctypes.byref(s)), "lpmx_plane_swe_solver_create")
        self._s = s

    def close(self):
        if getattr(self, "_s", None) and getattr(self.e, "_h", None):
            self.e._L.lpmx_plane_swe_solver_destroy(self._s)
        self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, passive, active, mask, layout=LAYOUT_RIGHT):
        P, A = _plane_swe_structs(passive, active, mask)
        self.e._check(self.e._L.lpmx_plane_swe_solver_set_state(self._s, ctypes.byref(P), ctypes.byref(A), layout,
                                                                self.np_, self.na), "lpmx_plane_swe_solver_set_state")

    def get_state(self, passive, active, layout=LAYOUT_RIGHT):
        P, A = _plane_swe_structs(passive, active, None)
        self.e._check(self.e._L.lpmx_plane_swe_solver_get_state(self._s, ctypes.byref(P), ctypes.byref(A), layout,
                                                                self.np_, self.na), "lpmx_plane_swe_solver_get_state")

    def init_direct_sums(self, do_velocity=True):
        self.e._check(self.e._L.lpmx_plane_swe_solver_init_direct_sums(self._s, int(bool(do_velocity))),
                      "lpmx_plane_swe_solver_init_direct_sums")

    def advance(self, dt, f0, beta, g, n_steps=1):
        self.e._check(self.e._L.lpmx_plane_swe_solver_advance(self._s, float(dt), float(f0), float(beta), float(g),
                                                              n_steps), "lpmx_plane_swe_solver_advance")

    def interactions_per_eval(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self.e._check(self.e._L.lpmx_plane_swe_solver_interactions_per_eval(self._s, ctypes.byref(a), ctypes.byref(b)),
                      "lpmx_plane_swe_solver_interactions_per_eval")
        return a.value, b.value
