#!/bin/bash
# One GPU-box visit: parity tests, smoke, contract bench, ncu launch list and one full capture of the pair-sum kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench ic2d_rk2"; timeout 900 python bench.py --stepper ic2d_rk2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_ic2d.json 2> $OUT/bench_ic2d.err; cat $OUT/bench_ic2d.json; tail -3 $OUT/bench_ic2d.err
echo "== bench swe_rk2"; timeout 900 python bench.py --stepper swe_rk2 --workload tc2_cubed7 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_swe.json 2> $OUT/bench_swe.err; cat $OUT/bench_swe.json; tail -3 $OUT/bench_swe.err
echo "== examples"; for e in bve_rotation sphere_rh54 sphere_gaussian_vortex sphere_swe_tc2; do ./examples/_build/$e -d 5 -dt 0.005 -tf 0.02 -n 4 2>&1 | tail -2; done > $OUT/examples.log 2>&1; cat $OUT/examples.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== quick sweep"; timeout 900 python tools/quick_bench.py > $OUT/quick_bench.log 2>&1; cat $OUT/quick_bench.log
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum -s 2 -c 2 -o $OUT/pair_sum python tools/ncu_probe.py 7 > $OUT/ncu_full.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum -s 1 -c 2 -o $OUT/pair_sum_ic2d python bench.py --stepper ic2d_rk2 --workload rh54_cubed6 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_ic2d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum -s 1 -c 1 -o $OUT/pair_sum_swe python bench.py --stepper swe_rk2 --workload tc2_cubed5 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_swe.log 2>&1
ls -la $OUT
