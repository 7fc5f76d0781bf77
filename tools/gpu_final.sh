#!/bin/bash
# Round-end check as the driver runs it: pytest -m gpu, smoke(), the contract bench (both arms), plus the Incompressible2DRK2
# stepper (what sphere_rh54 / sphere_gaussian_vortex actually step with) at icos-8.
TAG=${1:-r1ac}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
echo "== bench ic2d icos8"; timeout 300 python bench.py --stepper ic2d_rk2 --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_ic2d_icos8.json 2> $OUT/bench_ic2d_icos8.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ic2d_icos8.json
