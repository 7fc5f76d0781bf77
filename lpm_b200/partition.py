"""Host-side mirror of the engine's target sharding.

BVE / Incompressible2D solvers (lpm_b200/csrc/lpmx_steppers.cu: build_target_lists): only leaf faces are sources.  The leaf
faces (in index order) and the non-sources (vertices, then divided faces in index order) are each cut into `world` equal
slices; rank r owns slice r of both and evaluates them as two index lists per stage -- list A: its leaf faces (their packed
records are what the ranks exchange: the contiguous range [r L / W, (r+1) L / W) of the leaf-compacted array), list B: its
non-sources.  A is summed first and its records travel while B is summed; equal |A| and |B| on every rank keep the ranks in
step.  The SWE and planar solvers still shard the concatenated list (vertices then faces) into contiguous ranges
(`target_offsets(nv + nf, W)`).  The reference has no multi-device path: this is the new framework's
addition."""
import numpy as np


def target_offsets(n_targets, world):
    """world+1 offsets; rank r owns [off[r], off[r+1])."""
    return [(r * n_targets) // world for r in range(world + 1)]


def target_lists(n_verts, face_mask, rank, world):
    """(list A, list B) of global indices into the concatenated target list (vertices then faces) -- lpmx_local_targets."""
    face_mask = np.asarray(face_mask)
    faces = n_verts + np.arange(len(face_mask))
    lf = faces[face_mask == 0]
    ns = np.concatenate([np.arange(n_verts), faces[face_mask != 0]])
    a, b = target_offsets(len(lf), world), target_offsets(len(ns), world)
    return lf[a[rank]:a[rank + 1]].astype(np.int64), ns[b[rank]:b[rank + 1]].astype(np.int64)


def leaf_offsets(face_mask, world):
    """world+1 offsets into the leaf-compacted source array: rank r contributes the records [off[r], off[r+1])."""
    return target_offsets(int((np.asarray(face_mask) == 0).sum()), world)


def interactions_per_eval(n_verts, n_faces, n_leaves):
    """every target against every leaf minus each leaf's self pair."""
    return float(n_verts + n_faces) * n_leaves - n_leaves
