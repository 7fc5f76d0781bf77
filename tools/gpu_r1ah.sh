#!/bin/bash
# 1-GPU visit r1ah: multi-step IC2D calls evaluate psi for the last step only -- GPU suite, then the timing.
TAG=${1:-r1ah}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== full gpu suite"; timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
echo "== multistep timing"; timeout 120 python tools/quick_ic2d_multistep.py 2>&1 | tee $OUT/ic2d_multistep.txt
