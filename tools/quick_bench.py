"""Development probe (not the contract bench): time velocity evaluations at a few sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lpm_b200.api import Engine, PolyMesh2d, BVESolver
from lpm_b200 import gallery

e = Engine(0)
print("fp64 peak TF/s:", e.fp64_peak_tflops(), e.fp64_peak_tflops())
stream = torch.cuda.ExternalStream(e.stream())
for seed, depth in [("icos", 4), ("cubed", 6), ("cubed", 7), ("icos", 7), ("icos", 8)]:
    t0 = time.time()
    m = PolyMesh2d(seed, depth)
    tm = time.time() - t0
    f = gallery.RossbyHaurwitz54(); f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    s = BVESolver(e, m.n_verts, m.n_faces)
    s.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, m.face_area, m.face_mask)
    s.init_velocity(); e.sync()
    _, inter = s.interactions_per_eval()
    nsteps = 3 if depth < 8 else 1
    dt = 0.025 * m.appx_mesh_size() / 0.09045016
    s.advance(dt, 2 * np.pi, 1); e.sync()
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        s.advance(dt, 2 * np.pi, nsteps)
        b.record(stream)
    e.sync()
    ms = a.elapsed_time(b) / nsteps
    print(f"{seed}-{depth}: nv={m.n_verts} nf={m.n_faces} leaves={m.n_face_leaves} mesh {tm:.2f}s  "
          f"RK4 step {ms:.3f} ms  {4*inter/ms*1e-9:.2f} G-inter/s  alg {4*inter*24/ms*1e-9:.2f} TF/s", flush=True)
    s.close()
