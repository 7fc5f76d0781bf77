"""numpy restatement of the surface-Laplacian step of SWERK2 (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED against the reference: the arithmetic lives in Compadre 1.6.2 (tools/README.md:12-19), a third-party
dependency that is absent from /root/reference and from this image.  What is restated is (a) the reference's own call
sites -- gmls::Params defaults (src/lpm_compadre.hpp:23-60), gmls::Neighborhoods (src/lpm_compadre.cpp:75-112: k-nearest
search for min_neighbors points, window = eps_multiplier x distance to the k-th, then all points inside the window),
gmls::sphere_scalar_gmls (src/lpm_compadre.hpp:163-195: ScalarTaylorPolynomial, MANIFOLD, QR, PointSample, Power
weights) and GatherMeshData / ScatterMeshData (src/mesh/lpm_gather_mesh_data_impl.hpp:14-66,
src/mesh/lpm_scatter_mesh_data_impl.hpp) -- and (b) Compadre's published algorithm: weighted least squares in a local
tangent chart with a polynomial graph reconstruction of the manifold and the Laplace-Beltrami operator of that chart.
The least-squares path here is deliberately different from the product's (explicit neighbour lists from a kd-tree,
sqrt-weighted design matrix, numpy lstsq = LAPACK SVD) so that agreement checks the mathematics, not a shared bug.
Anchors: spherical harmonics (lap Y = -l(l+1) Y) and the TC2 closed form (examples/sphere_swe_tc2.cpp:243-244)."""
import math

import numpy as np


def params(order=3, eps_multiplier=2.0, weight_pwr=2.0, manifold_order=None, min_neighbors=None):
    """gmls::Params(order, dim) (src/lpm_compadre.hpp:49-60); min_neighbors = Compadre::GMLS::getNP(order, 2)."""
    return {"samples_order": order, "manifold_order": order if manifold_order is None else manifold_order,
            "eps_multiplier": eps_multiplier, "weight_pwr": weight_pwr,
            "min_neighbors": (order + 1) * (order + 2) // 2 if min_neighbors is None else min_neighbors}


def gather(vert_data, face_data, face_mask):
    """GatherMeshData: all vertices, then the leaf faces in face order (index n_verts + leaf_idx(f))."""
    leaf = np.asarray(face_mask) == 0
    return np.concatenate([np.asarray(vert_data), np.asarray(face_data)[leaf]], axis=0)


def scatter(gathered, n_verts, face_mask, face_out):
    """ScatterMeshData: vertices get gathered[:n_verts]; leaf faces get their entry; divided faces are not written."""
    leaf = np.asarray(face_mask) == 0
    face_out = np.array(face_out, copy=True)
    face_out[leaf] = gathered[n_verts:]
    return np.array(gathered[:n_verts], copy=True), face_out


def _basis(order, u, v):
    cols = []
    for n in range(order + 1):
        for ay in range(n + 1):
            ax = n - ay
            cols.append(u ** ax * v ** ay / (math.factorial(ax) * math.factorial(ay)))
    return np.stack(cols, axis=1)


def neighborhoods(xyz, p):
    """(eps, neighbour index lists): Neighborhoods(host_colloc_src_tgt_crds, params)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(xyz)
    dk, _ = tree.query(xyz, k=p["min_neighbors"])
    eps = np.where(dk[:, -1] > 0, dk[:, -1], 1e-14) * p["eps_multiplier"]
    lists = [np.array(sorted(j for j in tree.query_ball_point(xyz[i], eps[i]) if np.sum((xyz[j] - xyz[i]) ** 2) < eps[i] ** 2))
             for i in range(xyz.shape[0])]
    return eps, lists


def sphere_laplacian(xyz, f, p, targets=None):
    """Laplace-Beltrami of the samples f at the (collocated) points xyz.  Returns (lap, eps, n_neighbors)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    eps, lists = neighborhoods(xyz, p)
    n = xyz.shape[0]
    lap = np.full(n, np.nan)
    idx = range(n) if targets is None else targets
    for i in idx:
        x = xyz[i]
        nrm = x / np.linalg.norm(x)
        # any orthonormal tangent frame: take the two right-singular vectors of the projector's complement
        a = np.eye(3)[np.argmin(np.abs(nrm))]
        t1 = np.cross(nrm, a)
        t1 /= np.linalg.norm(t1)
        t2 = np.cross(nrm, t1)
        d = xyz[lists[i]] - x
        s, t, h = d @ t1, d @ t2, d @ nrm
        e = eps[i]
        w = np.maximum(1 - np.sqrt(s * s + t * t) / e, 0.0) ** p["weight_pwr"]
        sw = np.sqrt(w)
        Pf = _basis(p["samples_order"], s / e, t / e)
        Ph = _basis(p["manifold_order"], s / e, t / e)
        af = np.linalg.lstsq(Pf * sw[:, None], f[lists[i]] * sw, rcond=None)[0]
        ah = np.linalg.lstsq(Ph * sw[:, None], h * sw, rcond=None)[0]
        d2 = lambda c, k: (c[k] / e ** 2 if len(c) > 3 else 0.0)  # noqa: E731
        fs, ft, hs, ht = af[1] / e, af[2] / e, ah[1] / e, ah[2] / e
        g = 1 + hs * hs + ht * ht
        q = (hs * fs + ht * ft) / g
        ginv = np.array([[1 + ht * ht, -hs * ht], [-hs * ht, 1 + hs * hs]]) / g
        H = np.array([[d2(af, 3) - d2(ah, 3) * q, d2(af, 4) - d2(ah, 4) * q],
                      [d2(af, 4) - d2(ah, 4) * q, d2(af, 5) - d2(ah, 5) * q]])
        lap[i] = float((ginv * H).sum())
    return lap, eps, np.array([len(l) for l in lists])


def sphere_interpolate(src_xyz, src_fields, tgt_xyz, p):
    """ScalarPointEvaluation of CompadreRemesh (src/mesh/lpm_compadre_remesh_impl.hpp:136-210) with
    gmls::Neighborhoods(src, tgt, params): value at each target of the weighted least-squares Taylor fit of every field
    (rows of src_fields) over the source points inside the target's window."""
    from scipy.spatial import cKDTree
    src_xyz, tgt_xyz = np.asarray(src_xyz, dtype=np.float64), np.asarray(tgt_xyz, dtype=np.float64)
    F = np.atleast_2d(np.asarray(src_fields, dtype=np.float64))
    tree = cKDTree(src_xyz)
    dk, _ = tree.query(tgt_xyz, k=p["min_neighbors"])
    eps = np.where(dk[:, -1] > 0, dk[:, -1], 1e-14) * p["eps_multiplier"]
    out = np.full((F.shape[0], tgt_xyz.shape[0]), np.nan)
    for i, x in enumerate(tgt_xyz):
        nb = np.array([j for j in tree.query_ball_point(x, eps[i]) if np.sum((src_xyz[j] - x) ** 2) < eps[i] ** 2])
        nrm = x / np.linalg.norm(x)
        a = np.eye(3)[np.argmin(np.abs(nrm))]
        t1 = np.cross(nrm, a)
        t1 /= np.linalg.norm(t1)
        t2 = np.cross(nrm, t1)
        d = src_xyz[nb] - x
        s, t = d @ t1, d @ t2
        e = eps[i]
        sw = np.sqrt(np.maximum(1 - np.sqrt(s * s + t * t) / e, 0.0) ** p["weight_pwr"])
        P = _basis(p["samples_order"], s / e, t / e) * sw[:, None]
        coef = np.linalg.lstsq(P, (F[:, nb] * sw).T, rcond=None)[0]
        out[:, i] = coef[0]
    return out
