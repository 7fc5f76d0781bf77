"""Host-side mirror of the engine's target sharding (lpm_b200/csrc/lpmx_steppers.cu: solver_set_state).

The concatenated target list (vertices then faces) is split into `world` contiguous ranges; rank r evaluates
targets [t_r, t_{r+1}).  Only leaf faces are sources, so after every stage rank r owns the packed source
records of the leaves among its faces: the contiguous leaf range [l_r, l_{r+1}) given by faces.leaf_idx.
The reference has no multi-device path: this is the new framework's addition.
"""
import numpy as np


def target_offsets(n_targets, world):
    """world+1 offsets; rank r owns [off[r], off[r+1])."""
    return [(r * n_targets) // world for r in range(world + 1)]


def leaf_offsets(n_verts, face_mask, world):
    """world+1 offsets into the leaf-compacted source array matching target_offsets(n_verts + n_faces)."""
    face_mask = np.asarray(face_mask)
    n_faces = len(face_mask)
    leaf_idx = np.concatenate([[0], np.cumsum(face_mask == 0)]).astype(np.int64)  # exclusive scan, length nf+1
    out = []
    for t in target_offsets(n_verts + n_faces, world):
        f = min(max(t - n_verts, 0), n_faces)
        out.append(int(leaf_idx[f]))
    return out


def interactions_per_eval(n_verts, n_faces, n_leaves):
    """every target against every leaf minus each leaf's self pair."""
    return float(n_verts + n_faces) * n_leaves - n_leaves
