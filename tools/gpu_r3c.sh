#!/bin/bash
# r3c (8 GPUs): cubed-7 as the driver runs its scaling step (merged lists through the banks, 38 CTAs per launch and rank), and
# icos-8 (lists of 163 840 + 136 532 targets per rank: each through the banks by itself, exchange overlapped)
TAG=${1:-r3c}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
export LPMX_PEER_TIMEOUT_S=60
run() {
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%s: value %.4e  ms/step %.3f  e2e %.3f ms  launches %d bank %s  frac %.3f issued %.3f  parity %s  ic2d %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], r.get("bank_launches"), r["frac"], r["issued_frac"], (d.get("parity") or {}).get("max_rel_err"), (d.get("ic2d_rk2") or {}).get("ms_per_step")))
    print("  step_ms_each", [round(x, 3) for x in d["step_ms_each"][:8]])
except Exception as e:
    print(sys.argv[2], "no result:", e)
PY
  tail -2 $OUT/bench_$name.err | cut -c1-300
}
run cubed7_n$N X=0 -- --steps 20 --warmup 5
run icos8_n$N X=0 -- --workload gauss_icos8 --steps 3 --warmup 2 --no-extras
