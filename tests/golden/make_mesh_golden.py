"""Regenerate tests/golden/mesh_*.npz from the reference's seed files with the independent Python mesh
restatement (oracle/mesh_oracle.py).  Run in the build container (needs /root/reference/mesh_seeds):
    python tests/golden/make_mesh_golden.py
Also writes tests/golden/seed_tables.npz (the parsed .dat files) so the embedded seed tables can be checked
without the reference mount."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mesh_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("icos", 0), ("icos", 1), ("icos", 2), ("icos", 3), ("cubed", 0), ("cubed", 1), ("cubed", 2), ("cubed", 3),
         ("cubed", 4)]
# planar seeds (PlaneGeometry): (seed, depth, radius); radius 4 is the reference's own test (tests/lpm_polymesh_tests.cpp:54-67),
# radius 6 the default of examples/plane_gravity_wave.cpp
PLANE_CASES = [("quad_rect", 0, 1.0), ("quad_rect", 2, 1.0), ("quad_rect", 3, 4.0), ("quad_rect", 4, 6.0),
               ("tri_hex", 0, 1.0), ("tri_hex", 2, 1.0), ("tri_hex", 3, 6.0)]

if __name__ == "__main__":
    for seed, depth in CASES:
        m = mesh_oracle.TreeMesh(seed, depth)
        np.savez_compressed(os.path.join(HERE, f"mesh_{seed}_{depth}.npz"), **m.arrays())
        print(seed, depth, len(m.vx), len(m.eo), len(m.fx))
    for seed, depth, radius in PLANE_CASES:
        m = mesh_oracle.TreeMesh(seed, depth, radius=radius)
        np.savez_compressed(os.path.join(HERE, f"mesh_{seed}_{depth}_r{radius:g}.npz"), **m.arrays())
        print(seed, depth, radius, len(m.vx), len(m.eo), len(m.fx))
    tabs = {}
    for seed, d in mesh_oracle.SEEDS.items():
        crds, edges, fv, fe = mesh_oracle.read_seed(os.path.join("/root/reference/mesh_seeds", d["file"]), d["nverts"],
                                                    d["nfaces"], d["nedges"], d["nfv"], d.get("ndim", 3))
        tabs[f"{seed}_crds"] = np.array(crds)
        tabs[f"{seed}_edges"] = np.array(edges, dtype=np.int32)
        tabs[f"{seed}_face_verts"] = np.array(fv, dtype=np.int32)
        tabs[f"{seed}_face_edges"] = np.array(fe, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "seed_tables.npz"), **tabs)
