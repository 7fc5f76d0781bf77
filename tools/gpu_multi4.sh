#!/bin/bash
# 4-GPU visit: contract bench at N=4 and the synthetic sweep at N=4 (gpurun --gpus 4).
TAG=${1:-r1q}; N=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "== bench N=$N rc=$?"; cat $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/synthetic_sweep.py --sizes 1e4,3e4,1e5,3e5,1e6,3e6 --steps 1 > $OUT/sweep_n$N.jsonl 2> $OUT/sweep_n$N.err
echo "== sweep N=$N rc=$?"; cat $OUT/sweep_n$N.jsonl; tail -3 $OUT/sweep_n$N.err
