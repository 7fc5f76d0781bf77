"""Bank launches with and without programmatic dependent launch (LPMX_CONST_PDL), at cubed-7 on the resident solver, on the SAME
split of the targets (LPMX_CONST_SHAPE=6,4,3: 299 CTAs of 768 targets either way): the states after three BVERK4 steps must be
bit-identical (the reduction into the accumulators waits for the preceding launch, so the order of the additions per target is
the launch order with or without the overlap)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpm_b200 import gallery  # noqa: E402
from lpm_b200.api import BVESolver, Engine, PolyMesh2d  # noqa: E402

m = PolyMesh2d("cubed", 7)
f = gallery.RossbyHaurwitz54()
f.set_stationary_wave_speed()
vz, fz = f(m.vert_xyz), f(m.face_xyz)
area, mask = np.ascontiguousarray(m.face_area), np.ascontiguousarray(m.face_mask)
os.environ.setdefault("LPMX_CONST_SHAPE", "6,4,3")
os.environ["LPMX_CONST_STREAM"] = "1"
runs = []
for pdl in ("0", "1", "1"):
    os.environ["LPMX_CONST_PDL"] = pdl
    e = Engine(0)
    s = BVESolver(e, m.n_verts, m.n_faces)
    s.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
    s.init_velocity()
    c0 = e.const_stream_launch_count()
    for _ in range(3):
        s.advance(0.003, 2 * np.pi, 1)
    out = [np.empty((m.n_verts, 3)), np.empty(m.n_verts), np.empty((m.n_verts, 3)), np.empty((m.n_faces, 3)), np.empty(m.n_faces),
           np.empty((m.n_faces, 3))]
    s.get_state(*out)
    runs.append((e.const_stream_launch_count() - c0, out))
    s.close()
    e.close()
same = all(np.array_equal(runs[0][1][k], runs[i][1][k]) for k in range(6) for i in (1, 2))
print("bank launches", [r[0] for r in runs], "PDL == no PDL bitwise:", same, "finite:", bool(np.isfinite(runs[1][1][2]).all()))
sys.exit(0 if same and runs[0][0] > 0 else 1)
