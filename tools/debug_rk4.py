"""Development probe: BVERK4 in-place parity vs the oracle at several sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_b200.api import Engine, PolyMesh2d
from lpm_b200 import gallery
from oracle import oracle

e = Engine(0)
for seed, depth in [("cubed", 4), ("cubed", 5), ("icos", 5), ("cubed", 6)]:
    m = PolyMesh2d(seed, depth)
    f = gallery.SolidBodyRotation()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    leaf = m.face_mask == 0
    a = (m.face_xyz, fz, m.face_area, m.face_mask)
    vu = e.bve_velocity(m.vert_xyz, *a)
    fu = e.bve_velocity(None, *a, collocated=True)
    fu[~leaf] = 0
    got = [m.vert_xyz.copy(), vz.copy(), vu.copy(), m.face_xyz.copy(), fz.copy(), fu.copy()]
    ref = [x.copy() for x in got]
    for step in range(2):
        e.bve_rk4_step(0.01, 0.0, *got, m.face_area, m.face_mask, n_steps=1)
        psi = e.bve_streamfn(m.vert_xyz, got[3], got[4], m.face_area, m.face_mask)
        oracle.bve_rk4_step(0.01, 0.0, *ref, m.face_area, m.face_mask, n_steps=1)
        for name, g, r, sel in [("vx", got[0], ref[0], None), ("vu", got[2], ref[2], None), ("fx", got[3], ref[3], leaf), ("fu", got[5], ref[5], leaf)]:
            if sel is not None: g, r = g[sel], r[sel]
            d = np.abs(g - r).max(axis=1)
            bad = np.where(~(d <= 1e-9))[0]
            print(f"{seed}-{depth} step {step} {name}: max diff {np.nanmax(d):.3e} nan {np.isnan(g).sum()} bad {len(bad)} {bad[:8]}", flush=True)
