#!/bin/bash
# r3h (1 GPU): last look at the final tree: smoke + the default bench line with its extras
TAG=${1:-r3h}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 200 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%.4e inter/s  %.3f ms  e2e %.3f ms  frac %.3f issued %.3f parity %s  n1m %.4e  ic2d %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['issued_frac'], d['parity']['max_rel_err'], d['n1m']['interactions_per_s'], d['ic2d_rk2']['ms_per_step']))"
