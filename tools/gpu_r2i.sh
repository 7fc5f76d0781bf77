#!/bin/bash
# r2i (2 GPUs): the split evaluation (list A / list B, exchange overlapped with list B when the peer path is on): single-GPU
# regression (suite, bench, harness U2/U4), then the multi-GPU tests and the N = 2 bench with the NCCL and the peer exchange.
TAG=${1:-r2i}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tune"; timeout 300 ./tools/tune_pair_sum_r2g 229376 98304 r2 > $OUT/tune.txt 2>&1; head -4 $OUT/tune.txt
echo "== gpu suite"; timeout 2400 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
cp gpurun_out/multi_gpu_check_n*.log gpurun_out/peer_exchange_check_n2.log $OUT/ 2>/dev/null
echo "== bench N=1"; timeout 900 python bench.py --no-cpu-baseline --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-260 $OUT/bench_n1.json
export LPMX_PEER_TIMEOUT_S=10
for mode in 0 1; do
  LPMX_PEER_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29610+mode)) bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_peer$mode.json 2> $OUT/bench_n${N}_peer$mode.err
  echo "== bench N=$N LPMX_PEER_EXCHANGE=$mode"; python - "$OUT/bench_n${N}_peer$mode.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4e  ms/step %.3f  e2e %.4e (%.3f ms, h2d %d d2h %d)  exchange: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["exchange"]))
    print("parity", json.dumps(d["parity"])[120:420]); print("ic2d", d.get("ic2d_rk2")); print("launches", d["gpu_launches"], "kernel share", d["roofline"]["kernel_share_of_step"])
except Exception as e:
    print("no result:", e)
PY
  tail -3 $OUT/bench_n${N}_peer$mode.err
done
