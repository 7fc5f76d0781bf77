#!/bin/bash
# One GPU-box visit for the planar path: parity tests, the whole GPU suite, planar timings, ncu of the planar kernel.
TAG=${1:-r1p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest plane"; timeout 900 python -m pytest tests/test_gpu_parity_plane.py -q -m gpu > $OUT/pytest_plane.log 2>&1; echo "rc=$?"; tail -40 $OUT/pytest_plane.log
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== plane timings"; timeout 600 python tools/quick_bench_plane.py 64 128 256 512 > $OUT/quick_bench_plane.log 2>&1; cat $OUT/quick_bench_plane.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== ncu plane"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_sum -s 4 -c 2 -o $OUT/pair_sum_plane python tools/quick_bench_plane.py 256 > $OUT/ncu_plane.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_plane.log
ls -la $OUT
