#!/bin/bash
# 1-GPU visit r1t: moment-based GMLS kernel -- parity, timings, SWERK2 bench with the device Laplacian, ncu evidence.
TAG=${1:-r1t}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== gmls tests"; timeout 600 python -m pytest tests/test_gmls.py tests/test_ftle.py tests/test_gpu_parity_swe_rk2.py -q -m gpu --tb=short > $OUT/pytest_new.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_new.log; tail -15 $OUT/pytest_new.log
echo "== quick_gmls"; timeout 300 python tools/quick_gmls.py > $OUT/quick_gmls.log 2>&1; echo "rc=$?"; tail -20 $OUT/quick_gmls.log
echo "== bench swe frozen"; timeout 600 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_swe_frozen.json 2> $OUT/bench_swe_frozen.err; echo "rc=$?"; cut -c1-300 $OUT/bench_swe_frozen.json
echo "== bench swe gmls"; timeout 600 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --laplacian gmls --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_swe_gmls.json 2> $OUT/bench_swe_gmls.err; echo "rc=$?"; cut -c1-300 $OUT/bench_swe_gmls.json; tail -3 $OUT/bench_swe_gmls.err
echo "== example tc2 gmls"; ./examples/_build/sphere_swe_tc2 -d 5 -dt 0.005 -tf 0.02 -n 4 2>&1 | tail -5 | tee $OUT/example_tc2_gmls.log
./examples/_build/sphere_swe_tc2 -d 5 -dt 0.005 -tf 0.02 -n 4 -lap exact 2>&1 | tail -5 | tee $OUT/example_tc2_exact.log
echo "== ncu launch list swe+gmls"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_swe_gmls.csv python bench.py --stepper swe_rk2 --workload tc2_cubed7 --laplacian gmls --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full gmls kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmls_laplacian -s 1 -c 1 -o $OUT/gmls_laplacian python bench.py --stepper swe_rk2 --workload tc2_cubed7 --laplacian gmls --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_gmls.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_full_gmls.log
ls -la $OUT
