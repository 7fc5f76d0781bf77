"""A short run for ncu: a few velocity evaluations on cubed-sphere depth 6 (24 576 sources)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_b200.api import Engine, PolyMesh2d, BVESolver
from lpm_b200 import gallery
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
e = Engine(0)
m = PolyMesh2d("cubed", depth)
f = gallery.RossbyHaurwitz54(); f.set_stationary_wave_speed()
s = BVESolver(e, m.n_verts, m.n_faces)
s.set_state(m.vert_xyz, f(m.vert_xyz), None, m.face_xyz, f(m.face_xyz), None, m.face_area, m.face_mask)
s.init_velocity()
s.advance(0.01, 2 * np.pi, 2)
e.sync()
print("done", e.launch_count())
