#!/bin/bash
# r3a (1 GPU): 24 banks, half-bank launches for small target sets, merged lists: the bank-path tests, bitwise check, default
# bench, and the rank-size sweep in the automatic mode (what make_plan picks for a rank's target counts) against the ring kernel.
TAG=${1:-r3a}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_const_stream.py tests/test_gpu_parity_bve.py tests/test_gpu_parity_ic2d_swe.py -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_const.log
echo "== bitwise"; timeout 300 python tools/pdl_check.py 2>&1 | tail -2 | tee $OUT/pdl_check.txt
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-extras > $OUT/bench.json 2> $OUT/bench.err; cut -c1-200 $OUT/bench.json
python - <<'PY' | tee $OUT/auto_rank_sizes.txt
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from lpm_b200 import gallery
from lpm_b200.api import Engine, PolyMesh2d
m = PolyMesh2d("cubed", 7)
f = gallery.RossbyHaurwitz54(); f.set_stationary_wave_speed()
fz = f(m.face_xyz)
rng = np.random.default_rng(3)
for n in (12288, 16384, 24576, 28672, 32768, 57344, 114688, 229376):
    x = rng.standard_normal((n, 3)); x /= np.linalg.norm(x, axis=1)[:, None]
    row = []
    for mode in (0, -1):
        e = Engine(0)
        e.pair_sum_const_stream(mode)
        for _ in range(2): e.bve_velocity(x, m.face_xyz, fz, m.face_area, m.face_mask)
        e.profile_enable(True); e.profile_read()
        e.bve_velocity(x, m.face_xyz, fz, m.face_area, m.face_mask); e.sync()
        n_k, k_ms, pairs = e.profile_read()
        row.append((k_ms, e.const_stream_launch_count()))
        e.close()
    print("n_tgt %7d  ring %.3f ms   automatic %.3f ms (%d bank launches)  x%.3f" % (n, row[0][0], row[1][0], row[1][1], row[0][0] / row[1][0]))
PY
