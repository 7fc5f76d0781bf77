#!/bin/bash
# 1-GPU visit r1s: new FTLE / GMLS / gather-scatter parity tests first, then the whole GPU suite, probes and the contract bench.
TAG=${1:-r1s}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== new tests"; timeout 600 python -m pytest tests/test_ftle.py tests/test_gmls.py -q -m gpu --tb=short > $OUT/pytest_new.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_new.log; tail -25 $OUT/pytest_new.log
echo "== quick_gmls"; timeout 300 python tools/quick_gmls.py > $OUT/quick_gmls.log 2>&1; echo "rc=$?"; cat $OUT/quick_gmls.log | tail -20
echo "== sweep N=1"; timeout 300 python tools/synthetic_sweep.py --sizes 1e4,3e4,1e5,3e5,1e6 --steps 1 > $OUT/sweep_n1.jsonl 2> $OUT/sweep_n1.err; echo "rc=$?"; cat $OUT/sweep_n1.jsonl
echo "== full gpu suite"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== examples"; for e in sphere_rh54 sphere_gaussian_vortex; do ./examples/_build/$e -d 4 -tf 0.02 -n 2 2>&1 | tail -3; done > $OUT/examples.log 2>&1; cat $OUT/examples.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-600 $OUT/bench.json
