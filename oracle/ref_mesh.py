"""ctypes binding of oracle/_ref/liblpm_ref_mesh.so: the REFERENCE's own mesh classes (PolyMesh2d<Seed>::tree_init,
divide_flagged_faces), BVESphere + BVERK4::advance_timestep and Incompressible2D + Incompressible2DRK2, compiled in place from /root/reference/src against
oracle/kokkos_shim (oracle/ref_mesh_driver.cpp, `make -C oracle ref`).  TEST INFRASTRUCTURE ONLY: imported by the golden
generators under tests/golden/ and by the live comparisons in tests/ (skipped where the library is not built).  The library
travels to the GPU box with the snapshot; /root/reference itself is needed only to build it.  The reference's MeshSeed reads
mesh_seeds/*.dat at run time: where /root/reference is not mounted (the GPU box) the four files are rewritten from
tests/golden/seed_tables.npz (the same numbers, as the reference's parser read them; repr() round-trips doubles) into a temporary
directory that $LPM_ORACLE_SEED_DIR points the library at (oracle/kokkos_shim/LpmConfig.h)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "liblpm_ref_mesh.so")
SEED_DIR = "/root/reference/mesh_seeds"  # LPM_MESH_SEED_DIR of oracle/kokkos_shim/LpmConfig.h
SEED_ID = {"icos": 0, "cubed": 1, "quad_rect": 2, "tri_hex": 3}
SEED_FILE = {"icos": "icosTriSphereSeed.dat", "cubed": "cubedSphereSeed.dat", "quad_rect": "quadRectSeed.dat",
             "tri_hex": "triHexSeed.dat"}
SEED_TABLES = os.path.join(os.path.dirname(_HERE), "tests", "golden", "seed_tables.npz")

_lib = None


def available():
    return os.path.exists(LIB) and (os.path.isdir(SEED_DIR) or os.path.exists(SEED_TABLES))


def write_seed_files(dirname):
    """The four seed files in the layout MeshSeed<Seed>::read_file parses (src/mesh/lpm_mesh_seed.cpp:20-206): a header line,
    nverts + nfaces coordinate lines, then the blocks edgeO / faceverts / faceedges / vertEdges, each `n` lines after its keyword."""
    t = np.load(SEED_TABLES)
    for seed, fname in SEED_FILE.items():
        crds, edges = t[f"{seed}_crds"], t[f"{seed}_edges"]
        out = ["x   y" + ("   z" if crds.shape[1] == 3 else "")]
        out += ["  ".join(repr(float(v)) for v in row) for row in crds]
        out.append("edgeO      edgeD       edgeLeft        edgeRight")
        out += ["  ".join(str(int(v)) for v in row) for row in edges]
        out.append("faceverts")
        out += [" ".join(str(int(v)) for v in row) for row in t[f"{seed}_face_verts"]]
        out.append("faceedges")
        out += [" ".join(str(int(v)) for v in row) for row in t[f"{seed}_face_edges"]]
        out.append("vertEdges")
        out += [" ".join(str(int(v)) for v in row) for row in t[f"{seed}_vert_edges"]]
        with open(os.path.join(dirname, fname), "w") as f:
            f.write("\n".join(out) + "\n")


def _seed_dir():
    if os.path.isdir(SEED_DIR) and not os.environ.get("LPM_ORACLE_FORCE_REWRITTEN_SEEDS"):
        return SEED_DIR
    import tempfile
    d = tempfile.mkdtemp(prefix="lpm_seed_files_")
    write_seed_files(d)
    return d


def lib():
    global _lib
    if _lib is None:
        os.environ["LPM_ORACLE_SEED_DIR"] = _seed_dir()
        L = ctypes.CDLL(LIB)
        L.ref_mesh_create.restype = ctypes.c_void_p
        L.ref_mesh_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.ref_mesh_destroy.argtypes = [ctypes.c_void_p]
        L.ref_mesh_counts.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L.ref_mesh_get.argtypes = [ctypes.c_void_p] * 20
        L.ref_mesh_divide_flagged.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.ref_mesh_divide_flagged.restype = ctypes.c_int
        L.ref_mesh_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


class RefMesh:
    """PolyMesh2d<Seed>(PolyMeshParameters(depth, radius, amr_buffer, amr_limit)) of the reference."""

    def __init__(self, seed, depth, radius=1.0, amr_buffer=0, amr_limit=0):
        self.seed = seed
        self._h = lib().ref_mesh_create(SEED_ID[seed], depth, float(radius), amr_buffer, amr_limit)
        if not self._h:
            raise ValueError(f"unknown seed {seed}")

    def close(self):
        if self._h:
            lib().ref_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def counts(self):
        c = (ctypes.c_int * 9)()
        lib().ref_mesh_counts(self._h, c)
        return dict(zip(["n_verts", "n_edges", "n_faces", "nfv", "ndim", "n_leaves", "nmaxverts", "nmaxedges", "nmaxfaces"],
                        list(c)))

    def divide_flagged_faces(self, flags):
        """Returns (refine_count, outcome): outcome 0 all flagged faces divided, 1 not enough memory (nothing divided),
        2 level limit reached (the rest divided) -- what PolyMesh2d::divide_flagged_faces reports through its logger."""
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        rc = ctypes.c_int()
        out = lib().ref_mesh_divide_flagged(self._h, fl.ctypes.data_as(ctypes.c_void_p), len(fl), ctypes.byref(rc))
        return rc.value, out

    def arrays(self):
        """Same keys as tests/golden/mesh_*.npz (+ face_crd_idx)."""
        c = self.counts()
        nv, ne, nf, k, nd = c["n_verts"], c["n_edges"], c["n_faces"], c["nfv"], c["ndim"]
        i32 = np.int32
        a = dict(vert_xyz=np.zeros((nv, nd)), vert_lag_xyz=np.zeros((nv, nd)), edge_origs=np.zeros(ne, i32),
                 edge_dests=np.zeros(ne, i32), edge_lefts=np.zeros(ne, i32), edge_rights=np.zeros(ne, i32),
                 edge_parents=np.zeros(ne, i32), edge_kids=np.zeros((ne, 2), i32), face_xyz=np.zeros((nf, nd)),
                 face_lag_xyz=np.zeros((nf, nd)), face_area=np.zeros(nf), face_mask=np.zeros(nf, np.uint8),
                 face_verts=np.zeros((nf, k), i32), face_edges=np.zeros((nf, k), i32), face_parent=np.zeros(nf, i32),
                 face_kids=np.zeros((nf, 4), i32), face_level=np.zeros(nf, i32), face_leaf_idx=np.zeros(nf, i32),
                 face_crd_idx=np.zeros(nf, i32))
        lib().ref_mesh_get(self._h, *[v.ctypes.data_as(ctypes.c_void_p) for v in a.values()])
        return a


def bve_rk4_run(seed, depth, dt, omega, n_steps, vert_zeta, face_zeta, with_psi=False):
    """BVESphere<Seed>(depth) with the given relative vorticity -> init_velocity() -> n_steps x BVERK4::advance_timestep
    [-> init_stream_fn()], all the reference's code.  Returns dict(vert_xyz, vert_zeta, vert_vel, vert_psi, face_...)."""
    m = RefMesh(seed, depth)
    c = m.counts()
    m.close()
    nv, nf = c["n_verts"], c["n_faces"]
    vz = np.ascontiguousarray(vert_zeta, dtype=np.float64)
    fz = np.ascontiguousarray(face_zeta, dtype=np.float64)
    assert vz.shape == (nv,) and fz.shape == (nf,)
    out = dict(vert_xyz=np.zeros((nv, 3)), vert_zeta=np.zeros(nv), vert_vel=np.zeros((nv, 3)), vert_psi=np.zeros(nv),
               face_xyz=np.zeros((nf, 3)), face_zeta=np.zeros(nf), face_vel=np.zeros((nf, 3)), face_psi=np.zeros(nf))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    f = lib().ref_bve_rk4_run
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 10 + [ctypes.c_int]
    rc = f(SEED_ID[seed], depth, dt, omega, n_steps, p(vz), p(fz), p(out["vert_xyz"]), p(out["vert_zeta"]), p(out["vert_vel"]),
           p(out["vert_psi"]), p(out["face_xyz"]), p(out["face_zeta"]), p(out["face_vel"]), p(out["face_psi"]), int(with_psi))
    if rc != 0:
        raise RuntimeError("ref_bve_rk4_run failed")
    return out


def ic2d_rk2_run(seed, depth, dt, omega, eps, n_steps, vert_zeta, face_zeta):
    """Incompressible2D<Seed>(depth, CoriolisSphere(omega), eps) with the given relative vorticity -> init_direct_sums() ->
    n_steps x advance_timestep(Incompressible2DRK2), all the reference's code (oracle/ref_ic2d_driver.cpp).  Returns
    dict(vert_xyz, vert_zeta, vert_vel, vert_psi, face_...)."""
    m = RefMesh(seed, depth)
    c = m.counts()
    m.close()
    nv, nf = c["n_verts"], c["n_faces"]
    vz = np.ascontiguousarray(vert_zeta, dtype=np.float64)
    fz = np.ascontiguousarray(face_zeta, dtype=np.float64)
    assert vz.shape == (nv,) and fz.shape == (nf,)
    out = dict(vert_xyz=np.zeros((nv, 3)), vert_zeta=np.zeros(nv), vert_vel=np.zeros((nv, 3)), vert_psi=np.zeros(nv),
               face_xyz=np.zeros((nf, 3)), face_zeta=np.zeros(nf), face_vel=np.zeros((nf, 3)), face_psi=np.zeros(nf))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    f = lib().ref_ic2d_rk2_run
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 10
    rc = f(SEED_ID[seed], depth, dt, omega, eps, n_steps, p(vz), p(fz), p(out["vert_xyz"]), p(out["vert_zeta"]), p(out["vert_vel"]),
           p(out["vert_psi"]), p(out["face_xyz"]), p(out["face_zeta"]), p(out["face_vel"]), p(out["face_psi"]))
    if rc != 0:
        raise RuntimeError("ref_ic2d_rk2_run failed")
    return out
