#!/bin/bash
# 1-GPU visit r1aa: planar meshes end to end -- SWERK4 / IC2D-RK2 on the generated QuadRect / TriHex meshes against the oracle,
# the two planar example drivers, the whole GPU suite, and a larger gravity-wave run for timing.
TAG=${1:-r1aa}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_parity_plane.py tests/test_examples.py tests/test_amr.py -q -m gpu --tb=short > $OUT/pytest_new.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_new.log; tail -30 $OUT/pytest_new.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== plane examples"; ./examples/_build/plane_gravity_wave -d 6 -tf 0.5 -n 5 2>&1 | tail -4 | tee $OUT/example_plane.log
./examples/_build/plane_gravity_wave -s tri -d 5 -tf 0.5 -n 5 2>&1 | tail -3 | tee -a $OUT/example_plane.log
./examples/_build/plane_colliding_dipoles -d 6 -tf 0.1 -n 4 -amr 2 -c 0.2 -zv 0.3 2>&1 | tail -5 | tee -a $OUT/example_plane.log
