#!/bin/bash
TAG=${1:-r2n}; N=2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for mode in 1; do
  LPMX_PROFILE_DUMP=1 LPMX_PEER_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+mode)) tools/e2e_phases.py 2>&1 | grep "rank" | tee -a $OUT/e2e_phases_dump.txt
done
