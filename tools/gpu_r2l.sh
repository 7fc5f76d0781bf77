#!/bin/bash
# r2l (N GPUs): contract bench with the NCCL exchange in line and with the peer exchange overlapped; multi-GPU tests at N = 2
TAG=${1:-r2l}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
export LPMX_PEER_TIMEOUT_S=10
if [ "$N" = "2" ]; then
  echo "== multi-GPU tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -5 | tee $OUT/pytest_multi.log
  cp gpurun_out/multi_gpu_check_n*.log gpurun_out/peer_exchange_check_n2.log $OUT/ 2>/dev/null
fi
for mode in 0 1; do
  LPMX_PEER_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29610+mode)) bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_peer$mode.json 2> $OUT/bench_n${N}_peer$mode.err
  echo "== bench N=$N LPMX_PEER_EXCHANGE=$mode"; python - "$OUT/bench_n${N}_peer$mode.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4e  ms/step %.3f  e2e %.4e (%.3f ms, h2d %d d2h %d)  exchange: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["exchange"]))
    print("parity", json.dumps(d["parity"])[120:420]); print("ic2d", d.get("ic2d_rk2")); print("launches", d["gpu_launches"], "kernel share", d["roofline"]["kernel_share_of_step"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("no result:", e)
PY
  tail -2 $OUT/bench_n${N}_peer$mode.err
done
