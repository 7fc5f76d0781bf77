#!/bin/bash
# First GPU visit of the peer exchange (lpmx_peer.cu; written after round 1's GPU budget was spent, so unmeasured):
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_peer.sh r2a 2'       (then the same with 8)
# 1. bit-parity of the two exchanges + timing table, 2. the contract bench with each exchange.
TAG=${1:-r2a}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
export LPMX_PEER_TIMEOUT_S=10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/peer_exchange_check.py --time icos-4,cubed-6,cubed-7 --steps 8 > $OUT/peer_check_n$N.txt 2> $OUT/peer_check_n$N.err
echo "peer_exchange_check exit $?"; tail -20 $OUT/peer_check_n$N.txt; tail -5 $OUT/peer_check_n$N.err
for mode in 0 1; do
  LPMX_PEER_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29610+mode)) bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_peer$mode.json 2> $OUT/bench_n${N}_peer$mode.err
  echo "== bench N=$N LPMX_PEER_EXCHANGE=$mode"; cat $OUT/bench_n${N}_peer$mode.json; tail -3 $OUT/bench_n${N}_peer$mode.err
done
LPMX_TEST_PEER=1 timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -8 | tee $OUT/pytest_multi.log
