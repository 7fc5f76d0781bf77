#!/bin/bash
# r2c: reciprocal seed written in place (no MOV per pair) A/B in the tuning harness, then the full GPU suite + contract bench
# with the rebuilt library (and the unrolled FP64 peak probe).
TAG=${1:-r2c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in 0 1; do
  echo "== tune hold$v"; timeout 300 ./tools/tune_pair_sum_hold$v 229376 98304 r2 > $OUT/tune_hold$v.txt 2>&1; cat $OUT/tune_hold$v.txt
done
echo "== full gpu suite"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== bench ic2d"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline > $OUT/bench_ic2d.json 2> $OUT/bench_ic2d.err; cut -c1-300 $OUT/bench_ic2d.json
echo "== bench swe"; timeout 300 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --no-cpu-baseline --steps 3 > $OUT/bench_swe.json 2> $OUT/bench_swe.err; cut -c1-300 $OUT/bench_swe.json
