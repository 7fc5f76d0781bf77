"""Development probe: time the device-side surface Laplacian (gather -> grid sort -> GMLS -> scatter) and a SWERK2 step
with the built-in provider against the same step with a frozen Laplacian."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from lpm_b200 import gallery
from lpm_b200.api import Engine, PolyMesh2d, SWESolver, LAYOUT_LEFT

e = Engine(0)
stream = torch.cuda.ExternalStream(e.stream())
dev = torch.device("cuda", 0)
tc = gallery.SphereTestCase2()
for seed, depth in [("cubed", 5), ("cubed", 7), ("icos", 7), ("icos", 8)]:
    m = PolyMesh2d(seed, depth)
    leaf = m.face_mask == 0
    x = np.concatenate([m.vert_xyz, m.face_xyz[leaf]])
    f = tc.surface_exact(x)
    exact = tc.surface_laplacian_exact(x)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
    ft = torch.from_numpy(f).to(dev)
    for order in (3, 4):
        lap = e.gmls_sphere_laplacian(xt, ft, order, layout=LAYOUT_LEFT)
        e.sync()
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3):
                lap = e.gmls_sphere_laplacian(xt, ft, order, layout=LAYOUT_LEFT)
            b.record(stream)
        e.sync()
        err = np.abs(lap.cpu().numpy() - exact).max() / np.abs(exact).max()
        print(f"{seed}-{depth} n={x.shape[0]} order {order}: {a.elapsed_time(b)/3:.3f} ms per Laplacian, rel err vs TC2 closed form {err:.2e}", flush=True)
    if depth > 7:
        continue
    nv, nf = m.n_verts, m.n_faces
    vz, fz = tc.vorticity(m.vert_xyz), tc.vorticity(m.face_xyz)
    P = {"xyz": m.vert_xyz, "vort": vz, "div": np.zeros(nv), "depth": tc.surface(m.vert_xyz), "surf": tc.surface(m.vert_xyz),
         "bottom": np.zeros(nv), "laps": tc.surface_laplacian_exact(m.vert_xyz)}
    A = {"xyz": m.face_xyz, "vort": fz, "div": np.zeros(nf), "area": m.face_area, "mass": tc.surface(m.face_xyz) * m.face_area,
         "depth": tc.surface(m.face_xyz), "surf": tc.surface(m.face_xyz), "bottom": np.zeros(nf),
         "laps": tc.surface_laplacian_exact(m.face_xyz)}
    P = {k: np.ascontiguousarray(v) for k, v in P.items()}
    A = {k: np.ascontiguousarray(v) for k, v in A.items()}
    for label, prov in (("frozen laps", None), ("gmls provider order 4", e.gmls_provider(4))):
        s = SWESolver(e, nv, nf, eps=0.0)
        s.set_state(P, A, m.face_mask)
        s.init_direct_sums(True)
        dt = 0.01 * m.appx_mesh_size()
        s.advance(dt, 2 * np.pi, tc.g, prov, 1)
        e.sync()
        l0 = e.launch_count()
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            s.advance(dt, 2 * np.pi, tc.g, prov, 2)
            b.record(stream)
        e.sync()
        print(f"   SWERK2 step, {label}: {a.elapsed_time(b)/2:.3f} ms, {(e.launch_count()-l0)//2} launches/step", flush=True)
        s.close()
