#!/bin/bash
# 1-GPU visit r1y: adaptive refinement (flag kernels, AMR meshes through the pair sums, AMR example drivers), the whole GPU
# suite, smoke, and the contract bench with its reference arm as the driver runs them.
TAG=${1:-r1y}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== new tests"; timeout 600 python -m pytest tests/test_amr.py tests/test_examples.py -q -m gpu --tb=short > $OUT/pytest_new.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_new.log; tail -30 $OUT/pytest_new.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== amr examples"; for rs in direct indirect; do ./examples/_build/sphere_gaussian_vortex -d 5 -tf 0.02 -n 4 -amr 2 -c 0.1 -rm 2 -rs $rs 2>&1 | tail -6; done | tee $OUT/example_amr.log
./examples/_build/sphere_rh54 -d 5 -tf 0.02 -n 4 -amr 2 -c 0.5 2>&1 | tail -4 | tee -a $OUT/example_amr.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-1200 $OUT/bench.json; tail -3 $OUT/bench.err
