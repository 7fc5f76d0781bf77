"""Development probe (not the contract bench): time the planar SWERK4 stepper and the planar direct sums at a few
sizes.  usage: python tools/quick_bench_plane.py [n ...]   (n x n leaf panels)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import plane_cases
from lpm_b200 import api
from lpm_b200.api import Engine

e = Engine(0)
print("fp64 peak TF/s:", e.fp64_peak_tflops(), e.fp64_peak_tflops())
stream = torch.cuda.ExternalStream(e.stream())
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512]
for n in sizes:
    P, A, mask, h = plane_cases.quad_case(n=n, radius=6.0)
    pse = plane_cases.pse_eps_of(h)
    sol = api.PlaneSWESolver(e, P["xy"].shape[0], A["xy"].shape[0], 0.0, pse, api.TOPO_PLANAR_GAUSSIAN_MOUNTAIN)
    sol.set_state(P, A, mask)
    sol.init_direct_sums(True)
    e.sync()
    _, inter = sol.interactions_per_eval()
    dt = 0.1 * h
    sol.advance(dt, 0.0, 0.0, 1.0, 1)
    e.sync()
    nsteps = 3 if n <= 256 else 1
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        sol.advance(dt, 0.0, 0.0, 1.0, nsteps)
        b.record(stream)
    e.sync()
    ms = a.elapsed_time(b) / nsteps
    op = {"xy": np.empty_like(P["xy"]), "depth": np.empty_like(P["depth"])}
    oa = {"xy": np.empty_like(A["xy"]), "area": np.empty_like(A["area"])}
    sol.get_state(op, oa)
    ok = np.isfinite(op["xy"]).all() and np.isfinite(oa["area"]).all()
    print(f"plane n={n}: nv={P['xy'].shape[0]} nf={A['xy'].shape[0]} leaves={n*n}  SWERK4 step {ms:.3f} ms  "
          f"{4*inter/ms*1e-9:.3f} T-inter/s x1e-3 = {4*inter/ms*1e-6:.1f} G-inter/s  finite={ok}", flush=True)
    sol.close()
