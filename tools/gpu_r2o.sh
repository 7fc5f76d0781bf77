#!/bin/bash
# r2o (1 GPU): state at the end of round 2 as the driver will run it: suite with the error table, smoke, both bench arms,
# ncu launch list + one --set full capture of the default velocity kernel, IC2D / SWE stepper lines.
TAG=${1:-r2o}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -10 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_launches.log 2>&1
python tools/ncu_summarise.py launches $OUT/launches.csv > $OUT/launches.txt 2>&1; head -20 $OUT/launches.txt
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum_kernel -s 8 -c 1 -o $OUT/pair_sum_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_full.log 2>&1
python tools/ncu_summarise.py full $OUT/pair_sum_full.ncu-rep pair_sum > $OUT/pair_sum_ncu_full.txt 2>&1; head -30 $OUT/pair_sum_ncu_full.txt
echo "== ic2d / swe"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> /dev/null; cut -c1-200 $OUT/bench_ic2d.json
timeout 300 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --no-cpu-baseline --no-extras --steps 3 > $OUT/bench_swe.json 2> /dev/null; cut -c1-200 $OUT/bench_swe.json
