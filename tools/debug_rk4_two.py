import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_b200.api import Engine, PolyMesh2d, BVESolver, IC2DSolver
from lpm_b200 import gallery
from oracle import oracle
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
e = Engine(0)
m = PolyMesh2d("cubed", depth)
f = gallery.SolidBodyRotation()
vz, fz = f(m.vert_xyz), f(m.face_xyz)
leaf = m.face_mask == 0
a = (m.face_xyz, fz, m.face_area, m.face_mask)
vu = oracle.bve_velocity(m.vert_xyz, *a)
fu = oracle.bve_velocity(None, *a, collocated=True)
def rep(tag, got, ref):
    for name, g, r, sel in [("vx", got[0], ref[0], None), ("vu", got[2], ref[2], None), ("fx", got[3], ref[3], leaf), ("fu", got[5], ref[5], leaf)]:
        if sel is not None: g, r = g[sel], r[sel]
        d = np.abs(g - r).max(axis=1)
        print(f"{tag} {name}: max diff {np.nanmax(d):.3e} bad {(~(d <= 1e-9)).sum()}", flush=True)
# (2) dt = 0: every stage evaluates the same state
got = [m.vert_xyz.copy(), vz.copy(), vu.copy(), m.face_xyz.copy(), fz.copy(), fu.copy()]
ref = [x.copy() for x in got]
s = BVESolver(e, m.n_verts, m.n_faces)
s.set_state(*got, m.face_area, m.face_mask)
s.advance(0.0, 0.0, 1)
s.get_state(*got)
rep("dt=0", got, ref)
# tiny dt
for dt in (1e-6, 1e-3):
    got = [m.vert_xyz.copy(), vz.copy(), vu.copy(), m.face_xyz.copy(), fz.copy(), fu.copy()]
    ref = [x.copy() for x in got]
    s.set_state(*got, m.face_area, m.face_mask)
    s.advance(dt, 0.0, 1)
    s.get_state(*got)
    oracle.bve_rk4_step(dt, 0.0, *ref, m.face_area, m.face_mask, n_steps=1)
    rep(f"dt={dt}", got, ref)
# (1) IC2D RK2
pu, ppsi = oracle.ic2d_sums(m.vert_xyz, *a)
au, apsi = oracle.ic2d_sums(None, *a, targets_are_sources=True)
got = [m.vert_xyz.copy(), vz.copy(), pu.copy(), ppsi.copy(), m.face_xyz.copy(), fz.copy(), au.copy(), apsi.copy()]
ref = [x.copy() for x in got]
e.ic2d_rk2_step(0.01, 0.0, 0.0, *got, m.face_area, m.face_mask, n_steps=1)
oracle.ic2d_rk2_step(0.01, 0.0, 0.0, *ref, m.face_area, m.face_mask, n_steps=1)
for name, i, sel in [("px", 0, None), ("pu", 2, None), ("ax", 4, leaf), ("au", 6, leaf)]:
    g, r = got[i], ref[i]
    if sel is not None: g, r = g[sel], r[sel]
    d = np.abs(g - r).max(axis=1)
    print(f"ic2d {name}: max diff {np.nanmax(d):.3e} bad {(~(d <= 1e-9)).sum()}", flush=True)
