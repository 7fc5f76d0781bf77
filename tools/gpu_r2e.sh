#!/bin/bash
# r2e: full GPU suite with the adjudicated tolerances + the constant-bank tests, the contract bench in its new format (parity
# block, n1m, ic2d_rk2, icos-4 CPU example), the reference arm as torchrun would start it, and the constant-bank shapes at icos-8.
TAG=${1:-r2e}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PARITY_LOG=$PWD/$OUT/parity_errors.jsonl
rm -f $LPMX_PARITY_LOG
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log
unset LPMX_PARITY_LOG
echo "== bench reference (as a torchrun worker would see the environment)"; OMP_NUM_THREADS=1 timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json; tail -2 $OUT/bench.err
for shape in auto 8,8 5,8 6,12 6,8; do
  if [ $shape = auto ]; then unset LPMX_CONST_SHAPE; else export LPMX_CONST_SHAPE=$shape; fi
  timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_icos8_shape_$shape.json 2> $OUT/bench_icos8_shape_$shape.err
  echo "icos-8 const shape $shape: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_icos8_shape_$shape.json').read().strip().splitlines()[-1]); print('%.4e inter/s  %.1f ms  launches %d  parity %s' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['parity'].get('max_rel_err')))" 2>&1)"
done | tee $OUT/icos8_const_shapes.txt
unset LPMX_CONST_SHAPE
LPMX_CONST_STREAM=0 timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_icos8_ring.json 2> /dev/null
echo "icos-8 ring kernel: $(cut -c1-200 $OUT/bench_icos8_ring.json)"
