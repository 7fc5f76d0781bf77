"""Host-side mirror of the engine's target sharding.

BVE / Incompressible2D solvers (lpm_b200/csrc/lpmx_steppers.cu: solver_set_state): rank r owns the vertex rows
[r nv / W, (r+1) nv / W) and the face rows [r nf / W, (r+1) nf / W), and evaluates them as two index lists per stage --
list A: its leaf faces (the only particles that are sources; their packed records are what the ranks exchange, the contiguous
leaf range [l_r, l_{r+1}) given by faces.leaf_idx), list B: its vertices and divided faces.  A is summed first and its records
travel while B is summed.  The SWE and planar solvers still shard the concatenated list (vertices then faces) into contiguous
ranges (`target_offsets(nv + nf, W)`).  The reference has no multi-device path: this is the new
framework's addition."""
import numpy as np


def target_offsets(n_targets, world):
    """world+1 offsets; rank r owns [off[r], off[r+1])."""
    return [(r * n_targets) // world for r in range(world + 1)]


def local_rows(n_verts, n_faces, rank, world):
    """((v0, v1), (f0, f1)): the vertex rows and face rows rank `rank` owns (lpmx_local_rows)."""
    v, f = target_offsets(n_verts, world), target_offsets(n_faces, world)
    return (v[rank], v[rank + 1]), (f[rank], f[rank + 1])


def target_lists(n_verts, face_mask, rank, world):
    """(list A, list B) of global indices into the concatenated target list (vertices then faces): A = own leaf faces,
    B = own vertices followed by own divided faces."""
    face_mask = np.asarray(face_mask)
    (v0, v1), (f0, f1) = local_rows(n_verts, len(face_mask), rank, world)
    faces = np.arange(f0, f1)
    leaf = face_mask[f0:f1] == 0
    a = n_verts + faces[leaf]
    b = np.concatenate([np.arange(v0, v1), n_verts + faces[~leaf]])
    return a.astype(np.int64), b.astype(np.int64)


def leaf_offsets(face_mask, world):
    """world+1 offsets into the leaf-compacted source array: the leaves among each rank's face rows."""
    face_mask = np.asarray(face_mask)
    leaf_idx = np.concatenate([[0], np.cumsum(face_mask == 0)]).astype(np.int64)  # exclusive scan, length nf+1
    return [int(leaf_idx[f]) for f in target_offsets(len(face_mask), world)]


def interactions_per_eval(n_verts, n_faces, n_leaves):
    """every target against every leaf minus each leaf's self pair."""
    return float(n_verts + n_faces) * n_leaves - n_leaves
