#!/bin/bash
# 1-GPU visit r1u: GMLS interpolation / remesh parity, whole GPU suite, timing of the remesh interpolation.
TAG=${1:-r1u}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== new tests"; timeout 600 python -m pytest tests/test_gmls.py tests/test_examples.py -q -m gpu --tb=short > $OUT/pytest_new.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_new.log; tail -25 $OUT/pytest_new.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== remesh example timing"; for rs in direct indirect; do ./examples/_build/sphere_rh54 -d 6 -tf 0.02 -n 4 -rm 2 -rs $rs 2>&1 | tail -3; done | tee $OUT/example_remesh.log
