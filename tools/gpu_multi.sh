#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded parity, then the contract bench at 1..N GPUs.
TAG=${1:-r1m}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -15 | tee $OUT/pytest_multi.log
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  echo "== N=$n"; cat $OUT/bench_n$n.json; tail -3 $OUT/bench_n$n.err
  if [ $n -gt 1 ]; then
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 2 --warmup 3 --workload gauss_icos8 > $OUT/bench_icos8_n$n.json 2> $OUT/bench_icos8_n$n.err
    echo "== icos8 N=$n"; cat $OUT/bench_icos8_n$n.json; tail -3 $OUT/bench_icos8_n$n.err
  fi
done
