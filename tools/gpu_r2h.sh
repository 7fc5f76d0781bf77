#!/bin/bash
# r2h (2 GPUs): ring kernel with block flushes every 4 chunks (harness + bench), the multi-GPU tests (sharded steppers vs oracle,
# sharded host I/O, peer exchange == NCCL exchange bit for bit), peer-exchange timing table, contract bench at N = 2 with each
# exchange.
TAG=${1:-r2h}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== tune"; timeout 300 ./tools/tune_pair_sum_r2g 229376 98304 r2 > $OUT/tune.txt 2>&1; cat $OUT/tune.txt
echo "== bench N=1"; timeout 900 python bench.py --no-cpu-baseline --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-260 $OUT/bench_n1.json
echo "== multi-GPU tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -15 | tee $OUT/pytest_multi.log
cp gpurun_out/multi_gpu_check_n*.log gpurun_out/peer_exchange_check_n2.log $OUT/ 2>/dev/null
export LPMX_PEER_TIMEOUT_S=10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/peer_exchange_check.py --time icos-4,cubed-6,cubed-7 --steps 8 > $OUT/peer_check_n$N.txt 2> $OUT/peer_check_n$N.err
echo "peer_exchange_check exit $?"; tail -20 $OUT/peer_check_n$N.txt; tail -5 $OUT/peer_check_n$N.err
for mode in 0 1; do
  LPMX_PEER_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29610+mode)) bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_peer$mode.json 2> $OUT/bench_n${N}_peer$mode.err
  echo "== bench N=$N LPMX_PEER_EXCHANGE=$mode"; python - "$OUT/bench_n${N}_peer$mode.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4e  ms/step %.3f  e2e %.4e (%.3f ms, h2d %d d2h %d)  exchange: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["exchange"]))
    print("parity", json.dumps(d["parity"])[120:420]); print("ic2d", d.get("ic2d_rk2"))
except Exception as e:
    print("no result:", e)
PY
  tail -3 $OUT/bench_n${N}_peer$mode.err
done
