#!/bin/bash
# r2t (2 GPUs): the constant-bank path under the sharded solvers (index lists, overlapped peer exchange): multi-GPU tests,
# a forced run on cubed-6 (parity block = velocity of the advanced state against the reference arithmetic), icos-8 in the
# automatic mode, and the contract bench at cubed-7 (ring kernel per rank: 114 688 targets do not fill a wave).
TAG=${1:-r2t}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
export LPMX_PEER_TIMEOUT_S=20
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%s: value %.4e  ms/step %.3f  e2e %.3f ms  launches %d bank %s  frac %.3f  parity %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], r.get("bank_launches"), r["frac"], (d.get("parity") or {}).get("max_rel_err")))
except Exception as e:
    print(sys.argv[2], "no result:", e)
PY
  tail -2 $OUT/bench_$name.err | cut -c1-300
}
if [ "$N" = "2" ]; then
  echo "== multi-GPU tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -5 | tee $OUT/pytest_multi.log
  cp gpurun_out/multi_gpu_check_n*.log gpurun_out/peer_exchange_check_n2.log $OUT/ 2>/dev/null
  run cubed6_forced_const LPMX_CONST_STREAM=1 LPMX_CONST_MIN_TARGETS=1 -- --workload rh54_cubed6 --steps 5 --warmup 3 --no-extras
  run cubed6_ring LPMX_CONST_STREAM=0 -- --workload rh54_cubed6 --steps 5 --warmup 3 --no-extras
fi
run cubed7 LPMX_PEER_EXCHANGE=1 -- --steps 10 --warmup 3
run icos8 LPMX_PEER_EXCHANGE=1 -- --workload gauss_icos8 --steps 2 --warmup 1 --no-extras
