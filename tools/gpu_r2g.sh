#!/bin/bash
# r2g: two-level summation with the running totals in shared memory: ring-kernel speed back?  tuning harness, suite, benches.
TAG=${1:-r2g}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tune"; timeout 300 ./tools/tune_pair_sum_r2g 229376 98304 r2 > $OUT/tune.txt 2>&1; cat $OUT/tune.txt
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
for cs in auto 0; do
  if [ $cs = auto ]; then unset LPMX_CONST_STREAM; else export LPMX_CONST_STREAM=$cs; fi
  timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_icos8_cs$cs.json 2> $OUT/bench_icos8_cs$cs.err
  echo "icos-8 LPMX_CONST_STREAM=$cs: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_icos8_cs$cs.json').read().strip().splitlines()[-1]); print('%.4e inter/s  %.1f ms  launches %d  parity %s' % (d['value'], d['ms_per_step'], d['gpu_launches'], json.dumps(d['parity'])[120:520]))" 2>&1)"
done | tee $OUT/icos8_const.txt
unset LPMX_CONST_STREAM
echo "== ic2d"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> $OUT/bench_ic2d.err; cut -c1-260 $OUT/bench_ic2d.json
