#!/bin/bash
# 1-GPU visit r1ai: planar SWERK4 without the unobservable potentials -- GPU suite, then the planar timings.
TAG=${1:-r1ai}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== full gpu suite"; timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
echo "== plane timings"; timeout 100 python tools/quick_bench_plane.py > $OUT/plane_timings.txt 2>&1; tail -5 $OUT/plane_timings.txt
