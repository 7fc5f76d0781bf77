"""Regenerate tests/golden/ref_flags.npz: outputs of the REFERENCE's refinement-flag functors
(/root/reference/src/mesh/lpm_refinement_flags.hpp compiled in place -> oracle/_ref/liblpm_ref.so, see
oracle/ref_flags_driver.cpp) on seeded inputs.  Run in the build container after `make -C oracle ref`:
    python tests/golden/make_ref_flags_golden.py
Inputs are stored next to the outputs, so the fixture is self-contained on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refinement_oracle as ro  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(seed, rng):
    """An adaptively refined golden mesh (non-trivial mask, mixed levels) with seeded fields."""
    g = np.load(os.path.join(HERE, f"mesh_amr_{seed}_1_random.npz"))
    fx, vx, mask = g["face_xyz"], g["vert_xyz"], g["face_mask"]
    fz = np.sin(3 * fx[:, 0]) * np.cos(2 * fx[:, 2]) + 0.3 * fx[:, 1]
    vz = np.sin(3 * vx[:, 0]) * np.cos(2 * vx[:, 2]) + 0.3 * vx[:, 1]
    fz[mask != 0] = 40.0 * rng.standard_normal(int((mask != 0).sum()))  # divided faces carry stale values
    lag = vx + 0.04 * rng.standard_normal(vx.shape)
    return dict(face_mask=mask, face_vals=fz, area=g["face_area"], vert_vals=vz, face_verts=g["face_verts"], vert_lag=lag)


ARGS = {"scalar_max": ("face_vals",), "scalar_integral": ("face_vals", "area"),
        "scalar_variation": ("face_vals", "vert_vals", "face_verts"), "flow_map_variation": ("face_verts", "vert_lag")}

if __name__ == "__main__":
    L = ro.ref_lib()
    assert L is not None, "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20261017)
    out = {}
    for seed in ("icos", "cubed"):
        inp = inputs(seed, rng)
        for k, v in inp.items():
            out[f"{seed}_{k}"] = v
        n = inp["face_mask"].shape[0]
        for kind in ro.KINDS:
            arr = {k: inp[k] for k in ARGS[kind]}
            for tag, relative, rtol, start, end in (("rel", 1, 0.35, 0, n), ("abs", 0, 0.02, n // 4, n - 7)):
                flags, count, tol = ro.ref_iterate(L, kind, inp["face_mask"], rtol, relative, start, end, **arr)
                out[f"{seed}_{kind}_{tag}_flags"] = flags
                out[f"{seed}_{kind}_{tag}_meta"] = np.array([rtol, tol, count, start, end, relative], dtype=np.float64)
                print(seed, kind, tag, count, tol)
    np.savez_compressed(os.path.join(HERE, "ref_flags.npz"), **out)
