#!/bin/bash
# r2j (2 GPUs): per-launch timeline of the split evaluation on both ranks, NCCL in line vs peer overlapped
TAG=${1:-r2j}; N=2
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PEER_TIMEOUT_S=10 LPMX_PROFILE_DUMP=1
for mode in 0 1; do
  LPMX_PEER_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29610+mode)) bench.py --gpus $N --steps 3 --warmup 3 --no-extras --no-parity > $OUT/bench_peer$mode.json 2> $OUT/bench_peer$mode.err
  echo "== mode $mode"; cut -c1-200 $OUT/bench_peer$mode.json; grep "lpmx profile" $OUT/bench_peer$mode.err | head -60
done
