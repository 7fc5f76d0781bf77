"""What a rank of a multi-GPU run does per evaluation, on ONE GPU: n_tgt targets x the 98 304 leaf sources of cubed-7 through the
velocity pair sum, for the target counts a rank has at N = 1 .. 8 (whole lists and the two lists separately), ring kernel against
the pipelined bank path with different numbers of banks in rotation and of source records per launch.  CUDA-event time of the pair sum
only (lpmx_profile_enable), third call (the second one captures the graph)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpm_b200 import gallery  # noqa: E402
from lpm_b200.api import Engine, PolyMesh2d  # noqa: E402

m = PolyMesh2d("cubed", 7)
f = gallery.RossbyHaurwitz54()
f.set_stationary_wave_speed()
fz = f(m.face_xyz)
rng = np.random.default_rng(3)
sizes = [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "12288,16384,28672,57344,114688,229376".split(","))]
configs = [("ring", 0, {})]
for batch in (1280, 640, 320):
    for banks in (8, 12, 16, 24):
        configs.append((f"batch {batch} x{banks}", 1, {"LPMX_CONST_MIN_TARGETS": "1", "LPMX_CONST_BATCH": str(batch), "LPMX_CONST_BANKS": str(banks)}))
keys = ["LPMX_CONST_MIN_TARGETS", "LPMX_CONST_BATCH", "LPMX_CONST_BANKS"]
for n in sizes:
    x = rng.standard_normal((n, 3))
    x /= np.linalg.norm(x, axis=1)[:, None]
    ref = None
    for name, mode, env in configs:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        e = Engine(0)
        try:
            e.pair_sum_const_stream(mode)  # per handle (the environment's LPMX_CONST_STREAM is read once per process)
            for _ in range(2):
                u = e.bve_velocity(x, m.face_xyz, fz, m.face_area, m.face_mask)
            e.profile_enable(True)
            e.profile_read()
            u = e.bve_velocity(x, m.face_xyz, fz, m.face_area, m.face_mask)
            e.sync()
            n_k, k_ms, pairs = e.profile_read()
            if ref is None:
                ref = u
            err = float(np.abs(u - ref).max() / np.abs(ref).max())
            print(json.dumps({"n_tgt": n, "config": name, "ms": round(k_ms, 4), "pairs_per_s": round(n * 98304.0 / (k_ms * 1e-3), 1),
                              "bank_launches": e.const_stream_launch_count(), "rel_diff_vs_ring": err}), flush=True)
        finally:
            e.close()
