import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_b200.api import Engine, PolyMesh2d, BVESolver
from lpm_b200 import gallery
from oracle import oracle
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mode = sys.argv[2] if len(sys.argv) > 2 else "inplace"
e = Engine(0)
m = PolyMesh2d("cubed", depth)
f = gallery.SolidBodyRotation()
vz, fz = f(m.vert_xyz), f(m.face_xyz)
leaf = m.face_mask == 0
a = (m.face_xyz, fz, m.face_area, m.face_mask)
vu = oracle.bve_velocity(m.vert_xyz, *a)
fu = oracle.bve_velocity(None, *a, collocated=True)
got = [m.vert_xyz.copy(), vz.copy(), vu.copy(), m.face_xyz.copy(), fz.copy(), fu.copy()]
ref = [x.copy() for x in got]
if mode == "inplace":
    e.bve_rk4_step(0.01, 0.0, *got, m.face_area, m.face_mask, n_steps=1)
else:
    s = BVESolver(e, m.n_verts, m.n_faces)
    s.set_state(*got, m.face_area, m.face_mask)
    if mode == "initvel":
        s.init_velocity()
        s.get_state(*got)
        for name, g, r in [("vu", got[2], vu), ("fu", got[5], fu)]:
            d = np.abs(g - r).max(axis=1)
            print(name, "init_velocity diff", d.max(), "bad", (d > 1e-9).sum())
        sys.exit(0)
    s.advance(0.01, 0.0, 1)
    s.get_state(*got)
oracle.bve_rk4_step(0.01, 0.0, *ref, m.face_area, m.face_mask, n_steps=1)
for name, g, r, sel in [("vx", got[0], ref[0], None), ("vu", got[2], ref[2], None), ("fx", got[3], ref[3], leaf), ("fu", got[5], ref[5], leaf)]:
    if sel is not None: g, r = g[sel], r[sel]
    d = np.abs(g - r).max(axis=1)
    bad = np.where(~(d <= 1e-9))[0]
    print(f"{mode} cubed-{depth} {name}: max diff {np.nanmax(d):.3e} bad {len(bad)} {bad[:12]}", flush=True)
