#!/bin/bash
# r2v (1 GPU): programmatic dependent launch between the bank launches (LPMX_CONST_PDL) and the small-CTA shape that lets
# a CTA of the next launch share the SM with two of the running one (LPMX_CONST_SHAPE=6,4,2: 4 compute warps + prefetch warp).
TAG=${1:-r2v}
OUT=gpurun_out/$TAG; mkdir -p $OUT
line() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; print('%.4e inter/s  %.3f ms  e2e %.3f ms  launches %d  bank launches %s  frac %.3f issued %.3f  parity %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], r.get('bank_launches'), r['frac'], r['issued_frac'], (d.get('parity') or {}).get('max_rel_err')))" 2>&1; }
run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "$name [$*]: $(line $OUT/bench_$name.json)" | tee -a $OUT/ab.txt; }
echo "== bitwise"; for shp in "" "6,4,2"; do LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=$shp timeout 300 python tools/pdl_check.py 2>&1 | tail -2 | tee -a $OUT/pdl_check.txt; done
echo "== parity under PDL"; LPMX_CONST_PDL=1 timeout 600 python -m pytest tests/test_const_stream.py tests/test_gpu_parity_bve.py -q -m gpu -k 'const or cubed7' 2>&1 | tail -3 | tee $OUT/pytest_pdl.log
rm -f $OUT/ab.txt
run default LPMX_X=0
run default_pdl LPMX_CONST_PDL=1
run small LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=6,4,2
run small_pdl LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=6,4,2 LPMX_CONST_PDL=1
run small_pdl_nopf LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=6,4,2 LPMX_CONST_PDL=1 LPMX_CONST_PREFETCH=0
run small_pdl_nograph LPMX_CONST_STREAM=1 LPMX_CONST_SHAPE=6,4,2 LPMX_CONST_PDL=1 LPMX_CONST_GRAPH=0
