"""Regenerate tests/golden/ref_bve_rk4.npz and ref_ic2d_rk2.npz: outputs of the REFERENCE's own BVESphere<Seed> + BVERK4::advance_timestep
(/root/reference/src/lpm_bve_sphere_impl.hpp, lpm_bve_rk4_impl.hpp:63-167, lpm_bve_rk4.cpp) compiled in place against
oracle/kokkos_shim (oracle/ref_mesh_driver.cpp -> oracle/_ref/liblpm_ref_mesh.so).  Run in the build container:
    python tests/golden/make_ref_stepper_golden.py
ref_ic2d_rk2.npz: the same for Incompressible2D<Seed> + Incompressible2DRK2::advance_timestep_impl
(/root/reference/src/lpm_incompressible2d_impl.hpp, lpm_incompressible2d_rk2_impl.hpp:75-172; oracle/ref_ic2d_driver.cpp), cases
cubed3_rh54 (2 steps, eps 0), icos3_rh54 (3 steps, eps 0), cubed4_gauss (2 steps, eps 0.05): xyz, zeta, velocity, stream function.
ref_bve_rk4.npz cases (key prefix):
  icos3_rh54    icos depth 3, Rossby-Haurwitz 54, Omega 2 pi, dt 0.01, 3 steps: xyz, zeta, velocity, stream function
  cubed3_rh54   cubed depth 3, same fields, 2 steps
  icos4_rot_3   BASELINE configs[0] (examples/bve_rotation: solid-body rotation, Omega 0) at icos depth 4, dt 0.0025: 3 steps;
                xyz + zeta only
  icos4_rot_100 the same for 100 steps ("3 and 100 steps at icos-4")
                dt: the example's default 0.01 passes its own Courant check at icos-4 (2 pi dt / h = 0.9) but is in a singular
                regime of the scheme as coded: stage positions x + dt/2 u are not re-projected, |x| = 1.0005, so kappa - x.y of
                the nearest vertex/face pairs (h^2/6 = 8e-4) crosses zero; the compiled reference itself returns |u| = 48 in
                stage 2 and max | |x| - 1 | = 220 after 3 steps, 3e4 after 100, and its result moves by 1.5e-3 with the
                compiler's FMA contraction.  No tolerance is meaningful there; dt = 0.0025 keeps (dt |u| / 2)^2 = 6e-5 well
                below 8e-4.
The vorticity handed to the reference is the array stored here (`*_vert_zeta0`, `*_face_zeta0`, from lpm_b200/gallery.py, which
tests/test_gallery.py pins against the reference's gallery functors), so every engine under test starts from identical bits.
Rows of divided faces are stored as they came out (NaN / huge where a divided icosahedral face sits on its centre child: those
values depend on the compiler's FMA contraction, see tests/test_oracle_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpm_b200 import gallery  # noqa: E402
from lpm_b200.api import PolyMesh2d  # noqa: E402
from oracle import ref_mesh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("icos3_rh54", "icos", 3, "rh54", 2 * np.pi, 0.01, 3, True),
         ("cubed3_rh54", "cubed", 3, "rh54", 2 * np.pi, 0.01, 2, True),
         ("icos4_rot_3", "icos", 4, "rotation", 0.0, 0.0025, 3, False),
         ("icos4_rot_100", "icos", 4, "rotation", 0.0, 0.0025, 100, False)]


IC2D_CASES = [("cubed3_rh54", "cubed", 3, "rh54", 2 * np.pi, 0.01, 0.0, 2),
              ("icos3_rh54", "icos", 3, "rh54", 2 * np.pi, 0.01, 0.0, 3),
              ("cubed4_gauss", "cubed", 4, "gauss", 2 * np.pi, 0.01, 0.05, 2)]


def vorticity(kind):
    if kind == "gauss":
        return gallery.GaussianVortexSphere()
    if kind == "rh54":
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        return f
    return gallery.SolidBodyRotation()


if __name__ == "__main__":
    out = {}
    for name, seed, depth, kind, omega, dt, n_steps, full in CASES:
        m = PolyMesh2d(seed, depth)
        f = vorticity(kind)
        vz, fz = f(m.vert_xyz), f(m.face_xyz)
        r = ref_mesh.bve_rk4_run(seed, depth, dt, omega, n_steps, vz, fz, with_psi=full)
        out[f"{name}_params"] = np.array([depth, omega, dt, n_steps], dtype=np.float64)
        out[f"{name}_vert_zeta0"], out[f"{name}_face_zeta0"] = vz, fz
        keys = ["vert_xyz", "vert_zeta", "face_xyz", "face_zeta"] + (["vert_vel", "vert_psi", "face_vel", "face_psi"] if full else [])
        for k in keys:
            out[f"{name}_{k}"] = r[k]
        print(name, m.n_verts, m.n_faces, "max |x| - 1:", np.abs(np.linalg.norm(r["vert_xyz"], axis=1) - 1).max())
    np.savez_compressed(os.path.join(HERE, "ref_bve_rk4.npz"), **out)

    out = {}
    for name, seed, depth, kind, omega, dt, eps, n_steps in IC2D_CASES:
        m = PolyMesh2d(seed, depth)
        f = vorticity(kind)
        vz, fz = f(m.vert_xyz), f(m.face_xyz)
        r = ref_mesh.ic2d_rk2_run(seed, depth, dt, omega, eps, n_steps, vz, fz)
        out[f"{name}_params"] = np.array([depth, omega, dt, n_steps, eps], dtype=np.float64)
        out[f"{name}_vert_zeta0"], out[f"{name}_face_zeta0"] = vz, fz
        for k, v in r.items():
            out[f"{name}_{k}"] = v
        print(name, m.n_verts, m.n_faces, "max |x| - 1:", np.abs(np.linalg.norm(r["vert_xyz"], axis=1) - 1).max())
    np.savez_compressed(os.path.join(HERE, "ref_ic2d_rk2.npz"), **out)
