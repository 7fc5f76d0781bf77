#!/bin/bash
# r2s (1 GPU): state after the constant-bank rework as the driver will run it -- full suite with the error table, smoke, the
# default bench line, ncu launch list + one --set full capture of the bank kernel, IC2D stepper line, icos-8, and the
# synthetic sweep's top sizes (N = 3e6 as RK4 steps, N = 1e7 as one velocity evaluation: BASELINE configs[4]'s upper end).
TAG=${1:-r2s}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_launches.log 2>&1
python tools/ncu_summarise.py launches $OUT/launches.csv > $OUT/launches.txt 2>&1; head -12 $OUT/launches.txt
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum_const -s 100 -c 1 -o $OUT/const_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_full.log 2>&1
python tools/ncu_summarise.py full $OUT/const_full.ncu-rep pair_sum_const > $OUT/const_ncu_full.txt 2>&1; head -30 $OUT/const_ncu_full.txt
echo "== ic2d"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> /dev/null; cut -c1-200 $OUT/bench_ic2d.json
echo "== icos-8"; timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/bench_icos8.json 2> $OUT/bench_icos8.err; cut -c1-200 $OUT/bench_icos8.json
echo "== synthetic 3e6, 1e7"; timeout 900 python tools/synthetic_sweep.py --sizes 1e6,3e6,1e7 --steps 1 > $OUT/synthetic_sweep_n1.jsonl 2> $OUT/synthetic.err; cat $OUT/synthetic_sweep_n1.jsonl; tail -3 $OUT/synthetic.err
