"""Seeded FTLE inputs (test infrastructure): a reference configuration and a smoothly deformed physical one.

sphere_case: cubed-sphere mesh (quadrilateral faces, divided parents included), physical = reference pushed through
a smooth flow map and left slightly OFF the sphere (time-discretisation error, which ComputeFTLE normalises away for
the face centre only, mesh/lpm_ftle.hpp:96-98).  plane_case: an n x n panel lattice with its vertex connectivity in
the reference's counter-clockwise order starting at the top-left vertex (QuadRectSeed, mesh_seeds/quadRectSeed.dat)."""
import numpy as np


def _flow_sphere(x, amp, rng):
    y = x + amp * np.stack([np.sin(2 * x[:, 1]) * x[:, 2], x[:, 2] * x[:, 0], np.cos(3 * x[:, 0]) * x[:, 1]], axis=1)
    y /= np.linalg.norm(y, axis=1)[:, None]
    return y * (1 + 1e-6 * rng.standard_normal((x.shape[0], 1)))


def sphere_case(mesh, amp=0.15, seed=20261020):
    rng = np.random.default_rng(seed)
    return {"geom": 0, "vert_ref": mesh.vert_lag_xyz.copy(), "face_ref": mesh.face_lag_xyz.copy(),
            "vert_phys": _flow_sphere(mesh.vert_lag_xyz, amp, rng), "face_phys": _flow_sphere(mesh.face_lag_xyz, amp, rng),
            "face_verts": mesh.face_verts.astype(np.int32), "mask": mesh.face_mask.copy()}


def plane_case(n=12, radius=1.5, amp=0.2, seed=20261021):
    rng = np.random.default_rng(seed)
    # offset lattice: as coded the planar branch divides by |vertex 2| (see the quirk in oracle/lpm_oracle.c), which a
    # vertex at the origin would turn into 0/0
    g = np.linspace(-radius, radius, n + 1) + 0.0137
    vx, vy = np.meshgrid(g, g, indexing="ij")
    vert_ref = np.stack([vx.ravel(), vy.ravel()], axis=1)
    vid = lambda i, j: i * (n + 1) + j  # noqa: E731
    fv, fc = [], []
    for i in range(n):
        for j in range(n):
            fv.append([vid(i, j + 1), vid(i, j), vid(i + 1, j), vid(i + 1, j + 1)])  # top-left, then counter-clockwise
            fc.append([0.5 * (g[i] + g[i + 1]), 0.5 * (g[j] + g[j + 1])])
    face_ref = np.array(fc)
    mask = np.zeros(n * n, dtype=np.uint8)
    mask[rng.choice(n * n, n, replace=False)] = 1  # a few divided panels: must be skipped, output untouched

    def flow(x):
        return x + amp * np.stack([np.sin(1.3 * x[:, 1]) + 0.3 * x[:, 0] * x[:, 1], np.cos(0.9 * x[:, 0]) * x[:, 1]], axis=1)
    return {"geom": 1, "vert_ref": vert_ref, "face_ref": face_ref, "vert_phys": flow(vert_ref), "face_phys": flow(face_ref),
            "face_verts": np.array(fv, dtype=np.int32), "mask": mask}
