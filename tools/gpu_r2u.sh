#!/bin/bash
# r2u (8 GPUs): the two BASELINE configurations that need the whole box, with the constant-bank path under the sharded solver:
# configs[2] sphere_gaussian_vortex on icos-9 (9.6 M targets x 5.24 M leaf sources), one BVERK4 step after one warm-up step,
# and configs[4]'s upper end, the synthetic set with N = 1e7 (one velocity evaluation = 1e14 interactions).
TAG=${1:-r2u}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
export LPMX_PEER_TIMEOUT_S=120
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29628 bench.py --gpus $N --workload gauss_icos9 --steps 1 --warmup 1 --no-extras > $OUT/bench_icos9_n$N.json 2> $OUT/bench_icos9_n$N.err
echo "== bench icos9 N=$N rc=$?"; python - "$OUT/bench_icos9_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("icos-9: value %.4e  ms/step %.1f  e2e %.1f ms  launches %d bank %s  frac %.3f issued %.3f  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], r.get("bank_launches"), r["frac"], r["issued_frac"], (d.get("parity") or {}).get("max_rel_err")))
except Exception as e:
    print("no result:", e)
PY
tail -3 $OUT/bench_icos9_n$N.err | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29630 \
  tools/synthetic_sweep.py --sizes 3e6,1e7 --steps 1 --eval-only-above 4e6 --n-check 8 > $OUT/synthetic_sweep_n$N.jsonl 2> $OUT/synthetic_sweep_n$N.err
echo "== synthetic sweep rc=$?"; cut -c1-420 $OUT/synthetic_sweep_n$N.jsonl; tail -3 $OUT/synthetic_sweep_n$N.err | cut -c1-300
