#!/bin/bash
# r2p (1 GPU): the constant-bank path with two banks of 1 280 records, whole waves through the banks and the remainder through
# the ring kernel (automatic from one full wave per rank): parity, A/B at cubed-7 and icos-8, shapes, launch list, ncu of one
# bank launch.
TAG=${1:-r2p}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== parity"; timeout 900 python -m pytest tests/test_const_stream.py tests/test_gpu_parity_bve.py -q -m gpu -k 'const or cubed7' --durations=5 > $OUT/pytest_const.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_const.log; tail -12 $OUT/pytest_const.log
unset LPMX_PARITY_LOG
line() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; print('%.4e inter/s  %.3f ms  e2e %.3f ms  launches %d  bank launches %s  frac %.3f issued %.3f  parity %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], r.get('bank_launches'), r['frac'], r['issued_frac'], (d.get('parity') or {}).get('max_rel_err')))" 2>&1; }
echo "== cubed-7 A/B" | tee $OUT/ab.txt
for cs in auto 0; do
  if [ $cs = auto ]; then unset LPMX_CONST_STREAM; else export LPMX_CONST_STREAM=$cs; fi
  timeout 600 python bench.py --no-cpu-baseline --no-extras > $OUT/bench_cubed7_cs$cs.json 2> $OUT/bench_cubed7_cs$cs.err
  echo "cubed-7 LPMX_CONST_STREAM=$cs: $(line $OUT/bench_cubed7_cs$cs.json)" | tee -a $OUT/ab.txt
done
unset LPMX_CONST_STREAM
for shp in 3,8,2 4,12,1 3,16,1 5,8,1; do
  LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=$shp timeout 600 python bench.py --no-cpu-baseline --no-extras --no-parity > $OUT/bench_cubed7_shape_$shp.json 2> /dev/null
  echo "cubed-7 LPMX_CONST_SHAPE=$shp: $(line $OUT/bench_cubed7_shape_$shp.json)" | tee -a $OUT/ab.txt
done
LPMX_CONST_PREFETCH=128 timeout 600 python bench.py --no-cpu-baseline --no-extras --no-parity > $OUT/bench_cubed7_pf128.json 2> /dev/null
echo "cubed-7 LPMX_CONST_PREFETCH=128: $(line $OUT/bench_cubed7_pf128.json)" | tee -a $OUT/ab.txt
LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=3,8,2 LPMX_CONST_PREFETCH=128 timeout 600 python bench.py --no-cpu-baseline --no-extras --no-parity > $OUT/bench_cubed7_shape_382_pf128.json 2> /dev/null
echo "cubed-7 LPMX_CONST_SHAPE=3,8,2 PREFETCH=128: $(line $OUT/bench_cubed7_shape_382_pf128.json)" | tee -a $OUT/ab.txt
echo "== ic2d"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> $OUT/bench_ic2d.err; echo "ic2d_rk2 cubed-7: $(line $OUT/bench_ic2d.json)" | tee -a $OUT/ab.txt
echo "== icos-8"
timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/bench_icos8.json 2> $OUT/bench_icos8.err
echo "icos-8 auto: $(line $OUT/bench_icos8.json)" | tee -a $OUT/ab.txt
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_launches.log 2>&1
python tools/ncu_summarise.py launches $OUT/launches.csv > $OUT/launches.txt 2>&1; head -20 $OUT/launches.txt
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum_const -s 100 -c 1 -o $OUT/const_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_full.log 2>&1
python tools/ncu_summarise.py full $OUT/const_full.ncu-rep pair_sum_const > $OUT/const_ncu_full.txt 2>&1; head -32 $OUT/const_ncu_full.txt
