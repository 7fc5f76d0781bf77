#!/bin/bash
# r3b (4 GPUs): the bank path under the sharded solver on the small mesh (cubed-7: merged lists, 24 banks in rotation) at
# N = 4 and N = 2, against the ring kernel (LPMX_CONST_STREAM=0) -- parity block of every line = velocity of the advanced state
# against the reference arithmetic; multi-GPU check against the oracle.
TAG=${1:-r3b}; N=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
export LPMX_PEER_TIMEOUT_S=20
run() { # name, nproc, env..., -- bench args
  local name=$1; local np=$2; shift; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $np "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%s: value %.4e  ms/step %.3f  e2e %.3f ms  launches %d bank %s  frac %.3f issued %.3f  parity %s  ic2d %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], r.get("bank_launches"), r["frac"], r["issued_frac"], (d.get("parity") or {}).get("max_rel_err"), (d.get("ic2d_rk2") or {}).get("ms_per_step")))
except Exception as e:
    print(sys.argv[2], "no result:", e)
PY
  tail -2 $OUT/bench_$name.err | cut -c1-300
}
echo "== multi-GPU check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/multi_gpu_check_n$N.log 2>&1; echo "rc=$?"; tail -3 $OUT/multi_gpu_check_n$N.log | cut -c1-300
run cubed7_n${N} $N X=0 -- --steps 20 --warmup 5
run cubed7_n${N}_ring $N LPMX_CONST_STREAM=0 -- --steps 20 --warmup 5 --no-extras --no-parity
run cubed7_n2 2 X=0 -- --steps 20 --warmup 5
