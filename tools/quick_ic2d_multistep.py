"""Timing of Incompressible2DRK2 on cubed-sphere depth 7: 8 single-step calls (psi every step) against one 8-step call
(psi for the last step only).  Resident solver, CUDA-synchronised wall clock."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpm_b200 import gallery  # noqa: E402
from lpm_b200.api import Engine, IC2DSolver, PolyMesh2d  # noqa: E402

e = Engine(0)
m = PolyMesh2d("cubed", 7)
f = gallery.RossbyHaurwitz54()
f.set_stationary_wave_speed()
s = IC2DSolver(e, m.n_verts, m.n_faces, eps=0.0)
s.set_state(m.vert_xyz, f(m.vert_xyz), None, m.face_xyz, f(m.face_xyz), None, np.ascontiguousarray(m.face_area), np.ascontiguousarray(m.face_mask))
s.init_direct_sums()
dt, Omega, n = 0.003, 2 * np.pi, 8
s.advance(dt, Omega, 2)
e.sync()
t0 = time.perf_counter()
for _ in range(n):
    s.advance(dt, Omega, 1)
e.sync()
t1 = time.perf_counter()
s.advance(dt, Omega, n)
e.sync()
t2 = time.perf_counter()
print(f"cubed-7 Incompressible2DRK2: {n} x advance(1): {(t1 - t0) / n * 1e3:.2f} ms per step; advance({n}): {(t2 - t1) / n * 1e3:.2f} ms per step")
