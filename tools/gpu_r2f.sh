#!/bin/bash
# r2f: two-level summation + lazy psi + bve_solve: full GPU suite with the error table, contract bench, icos-8 with the
# constant-bank path (auto) and with the ring kernel, tuning harness for the ring kernel after the two-level change.
TAG=${1:-r2f}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -22 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
for cs in auto 0; do
  if [ $cs = auto ]; then unset LPMX_CONST_STREAM; else export LPMX_CONST_STREAM=$cs; fi
  timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_icos8_cs$cs.json 2> $OUT/bench_icos8_cs$cs.err
  echo "icos-8 LPMX_CONST_STREAM=$cs: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_icos8_cs$cs.json').read().strip().splitlines()[-1]); print('%.4e inter/s  %.1f ms  launches %d  parity %s' % (d['value'], d['ms_per_step'], d['gpu_launches'], json.dumps(d['parity'])[120:520]))" 2>&1)"
done | tee $OUT/icos8_const.txt
unset LPMX_CONST_STREAM
echo "== ic2d stepper lines"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> $OUT/bench_ic2d.err; cut -c1-260 $OUT/bench_ic2d.json
echo "== swe"; timeout 300 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --no-cpu-baseline --no-extras --steps 3 > $OUT/bench_swe.json 2> $OUT/bench_swe.err; cut -c1-260 $OUT/bench_swe.json
