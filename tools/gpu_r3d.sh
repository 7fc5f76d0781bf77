#!/bin/bash
# r3d (1 GPU): state at the end of round 2 as the driver will run it: full suite with the error table, smoke, both bench arms,
# ncu launch list + one --set full capture of the bank kernel, IC2D / SWE stepper lines.
TAG=${1:-r3d}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -10 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_launches.log 2>&1
python tools/ncu_summarise.py launches $OUT/launches.csv > $OUT/launches.txt 2>&1; head -12 $OUT/launches.txt
echo "== ic2d / swe"; timeout 300 python bench.py --stepper ic2d_rk2 --no-cpu-baseline --no-extras > $OUT/bench_ic2d.json 2> /dev/null; cut -c1-200 $OUT/bench_ic2d.json
timeout 300 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --no-cpu-baseline --no-extras --steps 3 > $OUT/bench_swe.json 2> /dev/null; cut -c1-200 $OUT/bench_swe.json
