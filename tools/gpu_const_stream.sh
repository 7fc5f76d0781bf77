#!/bin/bash
# First GPU visit of the constant-bank velocity path (lpmx_const_stream.cu; written after round 1's GPU budget was spent):
#   gpurun --timeout 1500 -- 'bash tools/gpu_const_stream.sh r2b'
# 1. parity (oracle + default kernel), 2. A/B of the contract bench (default / overlapped / serial copies),
# 3. launch-shape sweep at cubed-7, 4. icos-8 (where launch overhead no longer matters), 5. ncu of one launch.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; LPMX_TEST_CONST=1 timeout 600 python -m pytest tests/test_const_stream.py -q -m gpu --tb=short 2>&1 | tail -8 | tee $OUT/pytest_const.log
for mode in 0 1 2; do
  LPMX_CONST_STREAM=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cubed7_cs$mode.json 2> $OUT/bench_cubed7_cs$mode.err
  echo "== cubed-7 LPMX_CONST_STREAM=$mode"; python - "$OUT/bench_cubed7_cs$mode.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4e  ms/step %.3f  launches %d  frac %.3f  issued_frac %.3f" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"].get("issued_frac") or 0))
except Exception as e:
    print("no result:", e)
PY
  tail -2 $OUT/bench_cubed7_cs$mode.err
done
for shape in 6,8 6,9 6,10 5,10 5,12 7,8 7,9 8,8 4,12; do
  LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=$shape timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cubed7_shape_$shape.json 2> /dev/null
  echo "shape $shape: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_cubed7_shape_$shape.json').read().strip().splitlines()[-1]); print('%.4e inter/s  %.3f ms' % (d['value'], d['ms_per_step']))" 2>&1)"
done | tee $OUT/shape_sweep.txt
for mode in 0 1; do
  LPMX_CONST_STREAM=$mode timeout 600 python bench.py --workload gauss_icos8 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_icos8_cs$mode.json 2> $OUT/bench_icos8_cs$mode.err
  echo "== icos-8 LPMX_CONST_STREAM=$mode: $(cut -c1-160 $OUT/bench_icos8_cs$mode.json)"
done
LPMX_CONST_STREAM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum_const -s 400 -c 1 -o $OUT/const_stream_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
python tools/ncu_summarise.py full $OUT/const_stream_full.ncu-rep pair_sum_const > $OUT/const_stream_ncu_summary.txt 2>&1 || true
