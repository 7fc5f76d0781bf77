"""Seeded planar particle sets for the planar parity tests (test infrastructure).

`quad_case(n, radius)` mimics what PolyMesh2d<QuadRectSeed>::tree_init leaves behind after uniform refinement: an
(n+1)^2 vertex lattice, (n/2)^2 divided parent panels first (mask = 1, area = 0; their centres coincide with
vertices, as in the reference's quad divider, src/mesh/lpm_faces_impl.hpp:506-517) and then n^2 leaf panels.
Fields follow examples/plane_gravity_wave.cpp: PlanarGaussianSurfacePerturbation over PlanarGaussianMountain
(src/lpm_surface_gallery.hpp:41-88), plus a smooth vorticity dipole and a weak divergence so every kernel term is
exercised.  `jitter` perturbs the lattice (seeded) so no term vanishes by symmetry."""
import numpy as np


def surface_perturbation(xy):
    return 1.0 + 0.1 * np.exp(-(20 * (xy[:, 0] + 1.125) ** 2 + 5 * xy[:, 1] ** 2))


def gaussian_mountain(xy):
    return 0.8 * np.exp(-5.0 * (xy[:, 0] ** 2 + xy[:, 1] ** 2))


def vorticity(xy):
    return (3.0 * np.exp(-2.0 * ((xy[:, 0] - 0.6) ** 2 + (xy[:, 1] - 0.2) ** 2))
            - 2.5 * np.exp(-3.0 * ((xy[:, 0] + 0.5) ** 2 + (xy[:, 1] + 0.3) ** 2)))


def divergence(xy):
    return 0.2 * np.sin(0.7 * xy[:, 0]) * np.cos(0.9 * xy[:, 1])


def quad_case(n=16, radius=2.0, jitter=0.05, seed=20261019, topo=True, with_parents=True):
    assert n % 2 == 0
    rng = np.random.default_rng(seed)
    h = 2.0 * radius / n
    g = np.linspace(-radius, radius, n + 1)
    vx, vy = np.meshgrid(g, g, indexing="ij")
    vert_xy = np.stack([vx.ravel(), vy.ravel()], axis=1)
    c = -radius + h * (np.arange(n) + 0.5)
    fx, fy = np.meshgrid(c, c, indexing="ij")
    leaf_xy = np.stack([fx.ravel(), fy.ravel()], axis=1)
    leaf_area = np.full(n * n, h * h)
    if jitter:
        vert_xy = vert_xy + jitter * h * rng.uniform(-1, 1, vert_xy.shape)
        leaf_xy = leaf_xy + jitter * h * rng.uniform(-1, 1, leaf_xy.shape)
        leaf_area = leaf_area * (1 + 0.1 * rng.uniform(-1, 1, leaf_area.shape))
    if with_parents:
        pc = -radius + 2 * h * (np.arange(n // 2) + 0.5)
        px, py = np.meshgrid(pc, pc, indexing="ij")
        par_xy = np.stack([px.ravel(), py.ravel()], axis=1)
        face_xy = np.concatenate([par_xy, leaf_xy])
        area = np.concatenate([np.zeros(par_xy.shape[0]), leaf_area])
        mask = np.concatenate([np.ones(par_xy.shape[0], np.uint8), np.zeros(leaf_xy.shape[0], np.uint8)])
    else:
        face_xy, area, mask = leaf_xy, leaf_area, np.zeros(leaf_xy.shape[0], np.uint8)
    face_xy = np.ascontiguousarray(face_xy)
    bot_v = gaussian_mountain(vert_xy) if topo else np.zeros(vert_xy.shape[0])
    bot_f = gaussian_mountain(face_xy) if topo else np.zeros(face_xy.shape[0])
    surf_v, surf_f = surface_perturbation(vert_xy), surface_perturbation(face_xy)
    passive = {"xy": vert_xy, "vort": vorticity(vert_xy), "div": divergence(vert_xy), "depth": surf_v - bot_v,
               "surf": surf_v, "bottom": bot_v}
    active = {"xy": face_xy, "vort": vorticity(face_xy), "div": divergence(face_xy), "area": area,
              "mass": (surf_f - bot_f) * area, "depth": surf_f - bot_f, "surf": surf_f, "bottom": bot_f}
    return passive, active, mask, h


def pse_eps_of(h, power=11.0 / 20):
    """pse::PSEKernel<PlaneGeometry>::get_epsilon (src/lpm_pse.hpp:19-23)"""
    return h ** power
