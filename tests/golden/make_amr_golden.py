"""Regenerate tests/golden/mesh_amr_*.npz: adaptively refined meshes from the REFERENCE ITSELF --
PolyMesh2d<Seed>::divide_flagged_faces of /root/reference/src/mesh compiled in place (oracle/ref_mesh_driver.cpp ->
oracle/_ref/liblpm_ref_mesh.so, binding oracle/ref_mesh.py).  Run in the build container:
    python tests/golden/make_amr_golden.py
(Round 1 used the Python replay oracle/mesh_oracle.py; the compiled reference reproduces every array and every
(refine_count, outcome) of those files bit for bit, so the files did not change.)
Each file holds the flag arrays of every refinement pass (so a test can replay them through the product's generator), the
(refine_count, outcome) of every pass and all mesh arrays after the last pass.

Cases
  circ   : the AMR start-up loop of examples/sphere_gaussian_vortex.cpp:89-118 -- ScalarIntegralFlag on |zeta| A of the
           Gaussian vortex, relative tolerance fixed after the first pass, amr_limit passes over the faces added by the
           previous pass;
  random : seeded pseudo-random flags on all current leaves for 3 passes with amr_limit = 2 and a small amr_buffer: hits
           neighbours two levels apart, the level limit (outcome 2) and "not enough memory" (outcome 1, nothing divided).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_mesh, refinement_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def nmaxfaces(seed, lev):
    n0 = {"icos": 20, "cubed": 6, "quad_rect": 4, "tri_hex": 6}[seed]
    return sum(n0 * 4 ** k for k in range(lev + 1))  # MeshSeed::set_max_allocations (lpm_mesh_seed.cpp:266-279)


def gaussian_vortex(xyz):
    """GaussianVortexSphere defaults (src/lpm_vorticity_gallery.hpp:68-102): strength 4 pi, shape 4, centre lon 0 lat pi/20."""
    lat0 = np.pi / 20
    c = np.array([np.cos(lat0), 0.0, np.sin(lat0)])
    d2 = ((xyz - c) ** 2).sum(axis=1)
    return 4 * np.pi * np.exp(-16.0 * d2)


def run_case(seed, depth, kind, amr_buffer, amr_limit, passes):
    m = ref_mesh.RefMesh(seed, depth, 1.0, amr_buffer, amr_limit)
    nmax = nmaxfaces(seed, depth + amr_buffer)
    assert m.counts()["nmaxfaces"] == nmax
    out = {"seed": seed, "depth": depth, "amr_buffer": amr_buffer, "amr_limit": amr_limit, "nmaxfaces": nmax}
    rng = np.random.default_rng(20261017)
    start, tol = 0, None
    results = []
    for it in range(passes):
        a = m.arrays()
        n = a["face_mask"].shape[0]
        if kind == "circ":
            z = gaussian_vortex(a["face_xyz"])
            if tol is None:
                tol = 0.25 * refinement_oracle.flag_max("scalar_integral", a["face_mask"], face_vals=z, area=a["face_area"])
            flags, _ = refinement_oracle.iterate("scalar_integral", a["face_mask"], tol, start, n, face_vals=z, area=a["face_area"])
        else:
            flags = ((rng.random(n) < (0.35 if it < 2 else 0.9)) & (a["face_mask"] == 0)).astype(np.uint8)
        out[f"flags_{it}"] = flags
        results.append(m.divide_flagged_faces(flags))
        start = n
    out["results"] = np.array(results, dtype=np.int32)
    a = m.arrays()
    a.pop("face_crd_idx")
    out.update(a)
    m.close()
    return out


CASES = [("icos", 2, "circ", 2, 2, 2), ("cubed", 2, "circ", 2, 2, 2), ("icos", 1, "random", 2, 2, 4), ("cubed", 1, "random", 2, 2, 4),
         ("quad_rect", 1, "random", 2, 2, 4), ("tri_hex", 1, "random", 2, 2, 4)]

if __name__ == "__main__":
    for seed, depth, kind, buf, lim, passes in CASES:
        d = run_case(seed, depth, kind, buf, lim, passes)
        np.savez_compressed(os.path.join(HERE, f"mesh_amr_{seed}_{depth}_{kind}.npz"), **d)
        print(seed, depth, kind, d["results"].tolist(), d["face_mask"].shape[0], int(d["face_level"].max()))
