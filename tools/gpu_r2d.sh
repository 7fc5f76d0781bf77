#!/bin/bash
# r2d: the new parity tests (compiled-reference stepper fixtures, cubed-6 full, cubed-7 sampled, synthetic vs reference
# arithmetic) with the table of observed errors, and the contract bench with the corrected FP64 peak probe.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -40 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
