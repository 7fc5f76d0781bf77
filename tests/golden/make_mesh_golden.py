"""Regenerate tests/golden/mesh_*.npz from the REFERENCE ITSELF: PolyMesh2d<Seed>(PolyMeshParameters(depth, radius))
of /root/reference/src/mesh compiled in place (oracle/ref_mesh_driver.cpp -> oracle/_ref/liblpm_ref_mesh.so, `make -C oracle
ref`; binding oracle/ref_mesh.py).  Run in the build container (needs /root/reference):
    python tests/golden/make_mesh_golden.py
Round 1 generated these files with the Python replay oracle/mesh_oracle.py; the compiled reference reproduces every array
of every file bit for bit (integers and coordinates), so the files did not change when the generator did -- the replay is
now just a second witness (tests/test_mesh.py compares all three).
Also writes tests/golden/seed_tables.npz (the reference's mesh_seeds/*.dat as its own MeshSeed parsed them: the depth-0 mesh)
so the embedded seed tables can be checked without the reference mount."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_mesh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("icos", 0), ("icos", 1), ("icos", 2), ("icos", 3), ("cubed", 0), ("cubed", 1), ("cubed", 2), ("cubed", 3),
         ("cubed", 4)]
# planar seeds (PlaneGeometry): (seed, depth, radius); radius 4 is the reference's own test (tests/lpm_polymesh_tests.cpp:54-67),
# radius 6 the default of examples/plane_gravity_wave.cpp
PLANE_CASES = [("quad_rect", 0, 1.0), ("quad_rect", 2, 1.0), ("quad_rect", 3, 4.0), ("quad_rect", 4, 6.0),
               ("tri_hex", 0, 1.0), ("tri_hex", 2, 1.0), ("tri_hex", 3, 6.0)]
KEYS = ["vert_xyz", "vert_lag_xyz", "edge_origs", "edge_dests", "edge_lefts", "edge_rights", "edge_parents", "edge_kids",
        "face_xyz", "face_lag_xyz", "face_area", "face_mask", "face_verts", "face_edges", "face_parent", "face_kids",
        "face_level", "face_leaf_idx"]


def mesh_arrays(seed, depth, radius=1.0):
    m = ref_mesh.RefMesh(seed, depth, radius)
    a = m.arrays()
    m.close()
    return {k: a[k] for k in KEYS}


if __name__ == "__main__":
    for seed, depth in CASES:
        a = mesh_arrays(seed, depth)
        np.savez_compressed(os.path.join(HERE, f"mesh_{seed}_{depth}.npz"), **a)
        print(seed, depth, len(a["vert_xyz"]), len(a["edge_origs"]), len(a["face_xyz"]))
    for seed, depth, radius in PLANE_CASES:
        a = mesh_arrays(seed, depth, radius)
        np.savez_compressed(os.path.join(HERE, f"mesh_{seed}_{depth}_r{radius:g}.npz"), **a)
        print(seed, depth, radius, len(a["vert_xyz"]), len(a["edge_origs"]), len(a["face_xyz"]))
    tabs = {}
    for seed in ref_mesh.SEED_ID:
        a = mesh_arrays(seed, 0)  # the seed as MeshSeed<Seed>::read_file parsed it (radius 1)
        tabs[f"{seed}_crds"] = np.concatenate([a["vert_xyz"], a["face_xyz"]])
        tabs[f"{seed}_edges"] = np.stack([a["edge_origs"], a["edge_dests"], a["edge_lefts"], a["edge_rights"]], axis=1).astype(np.int32)
        tabs[f"{seed}_face_verts"] = a["face_verts"].astype(np.int32)
        tabs[f"{seed}_face_edges"] = a["face_edges"].astype(np.int32)
    # the vertEdges block of each file (read by MeshSeed::read_file, not exposed by the mesh classes) straight from the .dat,
    # so that oracle/ref_mesh.write_seed_files can rewrite complete seed files where /root/reference is not mounted
    for seed, fname in ref_mesh.SEED_FILE.items():
        lines = open(os.path.join(ref_mesh.SEED_DIR, fname)).read().splitlines()
        k = next(i for i, ln in enumerate(lines) if "vertEdges" in ln)
        nv = len(tabs[f"{seed}_crds"]) - len(tabs[f"{seed}_face_verts"])
        tabs[f"{seed}_vert_edges"] = np.array([[int(t) for t in ln.split()] for ln in lines[k + 1:k + 1 + nv]], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "seed_tables.npz"), **tabs)
