#!/bin/bash
# 1-GPU visit r1ad: the 1024-entry / exponent-table log in the stream-function kernels -- whole GPU suite, then the steppers that use it.
TAG=${1:-r1ad}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
echo "== bench ic2d cubed7"; timeout 300 python bench.py --stepper ic2d_rk2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_ic2d_rk2.json 2> $OUT/bench_ic2d_rk2.err; echo "rc=$?"; cut -c1-260 $OUT/bench_ic2d_rk2.json; tail -2 $OUT/bench_ic2d_rk2.err
echo "== plane timings"; timeout 300 python tools/quick_bench_plane.py > $OUT/plane_timings.txt 2>&1; tail -12 $OUT/plane_timings.txt
