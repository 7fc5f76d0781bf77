#!/bin/bash
# r2w (1 GPU): pipelined bank launches (programmatic dependent launch) as the default: every target through the banks, CTAs of
# T = 6 x 4 compute warps + prefetch warp, three per SM.  Parity, bitwise PDL == no PDL, A/B at cubed-7 and icos-8, ncu.
TAG=${1:-r2w}
OUT=gpurun_out/$TAG; mkdir -p $OUT
line() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; print('%.4e inter/s  %.3f ms  e2e %.3f ms  launches %d  bank launches %s  frac %.3f issued %.3f  parity %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], r.get('bank_launches'), r['frac'], r['issued_frac'], (d.get('parity') or {}).get('max_rel_err')))" 2>&1; }
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift; env "${envs[@]}" timeout 400 python bench.py --no-cpu-baseline --no-extras "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "$name [${envs[*]}]: $(line $OUT/bench_$name.json)" | tee -a $OUT/ab.txt; }
echo "== parity"; timeout 900 python -m pytest tests/test_const_stream.py tests/test_gpu_parity_bve.py -q -m gpu -k 'const or cubed7 or split' 2>&1 | tail -3 | tee $OUT/pytest_const.log
echo "== bitwise"; LPMX_CONST_STREAM=1 timeout 300 python tools/pdl_check.py 2>&1 | tail -2 | tee $OUT/pdl_check.txt
rm -f $OUT/ab.txt
run cubed7_default X=0 --
run cubed7_pdl0 LPMX_CONST_PDL=0 -- --no-parity
run cubed7_T5 LPMX_CONST_SHAPE=5,4,3 -- --no-parity
run ic2d X=0 -- --stepper ic2d_rk2
run icos8_default X=0 -- --workload gauss_icos8 --steps 1 --warmup 1
run icos8_582 LPMX_CONST_SHAPE=5,8,2 -- --workload gauss_icos8 --steps 1 --warmup 1 --no-parity
run icos8_682 LPMX_CONST_SHAPE=6,8,2 -- --workload gauss_icos8 --steps 1 --warmup 1 --no-parity
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_sum_const -s 100 -c 1 -o $OUT/const_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity > $OUT/ncu_full.log 2>&1
python tools/ncu_summarise.py full $OUT/const_full.ncu-rep pair_sum_const > $OUT/const_ncu_full.txt 2>&1; head -30 $OUT/const_ncu_full.txt
