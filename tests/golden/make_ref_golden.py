"""Golden outputs of the REFERENCE's own functors (compiled in place: oracle/_ref/liblpm_ref.so, built by
`make -C oracle ref` where /root/reference is mounted) on small meshes -> tests/golden/ref_sums.npz.
    python tests/golden/make_ref_golden.py
The fixtures let any machine (incl. the GPU box, which has no /root/reference) check the oracle restatement
against real reference output."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpm_b200 import gallery  # noqa: E402
from lpm_b200.api import PolyMesh2d  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    R = ctypes.CDLL(O.REF_LIB)
    out = {}
    for seed, depth in (("icos", 2), ("cubed", 3)):
        m = PolyMesh2d(seed, depth)
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        fz = f(m.face_xyz)
        sig = 0.3 * m.face_xyz[:, 0] * m.face_xyz[:, 2]
        k = f"{seed}{depth}_"
        out[k + "zeta"] = fz
        out[k + "sigma"] = sig
        a = (m.face_xyz, fz, m.face_area, m.face_mask)
        out[k + "bve_vel_verts"] = O.bve_velocity(m.vert_xyz, *a, L=R)
        out[k + "bve_vel_faces"] = O.bve_velocity(None, *a, collocated=True, L=R)
        out[k + "bve_psi_verts"] = O.bve_streamfn(m.vert_xyz, *a, L=R)
        out[k + "bve_psi_faces"] = O.bve_streamfn(None, *a, collocated=True, L=R)
        for eps in (0.0, 0.05):
            e = f"eps{eps}_"
            u, p = O.ic2d_sums(m.vert_xyz, *a, eps=eps, L=R)
            out[k + e + "ic2d_vel_passive"], out[k + e + "ic2d_psi_passive"] = u, p
            u, p = O.ic2d_sums(None, *a, eps=eps, targets_are_sources=True, L=R)
            out[k + e + "ic2d_vel_active"], out[k + e + "ic2d_psi_active"] = u, p
            _, dd, _ = O.swe_sphere_sums(m.vert_xyz, m.face_xyz, fz, sig, m.face_area, m.face_mask, eps=eps, L=R)
            out[k + e + "swe_ddot_verts"] = dd
            _, dd, _ = O.swe_sphere_sums(None, m.face_xyz, fz, sig, m.face_area, m.face_mask, eps=eps,
                                         targets_are_sources=True, L=R)
            out[k + e + "swe_ddot_faces"] = dd
    # O(N) SWE functors (tendencies) on seeded random particle data
    rng = np.random.default_rng(20261018)
    n = 257
    x = rng.standard_normal((n, 3))
    x /= np.linalg.norm(x, axis=1)[:, None]
    x *= 1 + 1e-3 * rng.standard_normal((n, 1))
    tin = {"x": x, "u": rng.standard_normal((n, 3)), "zeta": rng.standard_normal(n), "sigma": rng.standard_normal(n),
           "third": 1 + rng.random(n), "ddot": rng.standard_normal(n), "laps": rng.standard_normal(n)}
    for k2, v in tin.items():
        out["tend_in_" + k2] = v
    for is_area in (0, 1):
        dz, ds, d3 = O.swe_tendencies(is_area, tin["x"], tin["u"], tin["zeta"], tin["sigma"], tin["third"], tin["ddot"],
                                      tin["laps"], Omega=2 * np.pi, g=1.5, dt=0.0125, L=R)
        out[f"tend_out_{is_area}"] = np.stack([dz, ds, d3])
    # pair-level values of the reference functions on random pairs, on and off the unit sphere
    rng = np.random.default_rng(20261017)
    xs, ys, epss, vals = [], [], [], []
    for it in range(64):
        x = rng.standard_normal(3)
        y = rng.standard_normal(3)
        x /= np.linalg.norm(x)
        y /= np.linalg.norm(y)
        if it % 3 == 1:
            x *= 1 + 1e-2 * rng.standard_normal()
            y *= 1 + 1e-2 * rng.standard_normal()
        eps = (0.0, 0.01, 0.1)[it % 3]
        kz, ks, gkz, gks = O.swe_pair(x, y, eps, L=R)
        xs.append(x), ys.append(y), epss.append(eps), vals.append(np.concatenate([kz, ks, gkz, gks]))
    out["pair_x"], out["pair_y"], out["pair_eps"], out["pair_vals"] = map(np.array, (xs, ys, epss, vals))
    np.savez_compressed(os.path.join(HERE, "ref_sums.npz"), **out)
    print("wrote", len(out), "arrays")
