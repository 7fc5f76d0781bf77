"""Development probe: operator-level parity against the oracle over a range of mesh sizes (finds size-dependent bugs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_b200.api import Engine, PolyMesh2d
from lpm_b200 import gallery
from oracle import oracle

e = Engine(0)
for seed, depth in [("cubed", 4), ("cubed", 5), ("icos", 4), ("icos", 5), ("cubed", 6)]:
    m = PolyMesh2d(seed, depth)
    f = gallery.SolidBodyRotation()
    fz = f(m.face_xyz)
    leaf = m.face_mask == 0
    a = (m.face_xyz, fz, m.face_area, m.face_mask)
    for name, got, ref, sel in [
        ("vel verts", e.bve_velocity(m.vert_xyz, *a), oracle.bve_velocity(m.vert_xyz, *a), None),
        ("vel faces", e.bve_velocity(None, *a, collocated=True), oracle.bve_velocity(None, *a, collocated=True), leaf),
        ("psi verts", e.bve_streamfn(m.vert_xyz, *a), oracle.bve_streamfn(m.vert_xyz, *a), None),
        ("psi faces", e.bve_streamfn(None, *a, collocated=True), oracle.bve_streamfn(None, *a, collocated=True), leaf),
    ]:
        if sel is not None:
            got, ref = got[sel], ref[sel]
        d = np.abs(got - ref)
        d = d if d.ndim == 1 else d.max(axis=1)
        bad = np.where(~(d <= 1e-10 * np.abs(ref).max()))[0]
        print(f"{seed}-{depth} {name}: max diff {np.nanmax(d):.3e} nan {np.isnan(got).sum()} bad {len(bad)} {bad[:10]}", flush=True)
