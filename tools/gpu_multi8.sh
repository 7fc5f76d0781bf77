#!/bin/bash
# 8-GPU visit: contract bench at N=8 (default workload and icos-8) and the synthetic sweep at N=8 (gpurun --gpus 8).
TAG=${1:-r1w}; N=8
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "== bench N=$N rc=$?"; cut -c1-400 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 tools/synthetic_sweep.py --sizes 1e5,3e5,1e6,3e6 --steps 1 > $OUT/sweep_n$N.jsonl 2> $OUT/sweep_n$N.err
echo "== sweep N=$N rc=$?"; cut -c1-330 $OUT/sweep_n$N.jsonl; tail -3 $OUT/sweep_n$N.err
