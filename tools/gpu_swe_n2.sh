#!/bin/bash
# 2-GPU visit: SWERK2 with the device GMLS Laplacian, target-sharded over 2 GPUs (gpurun --gpus 2): cubed-7 (contract-style line) and
# BASELINE configs[3] at its stated size (icos-8).
TAG=${1:-r1ab}; N=2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 bench.py --gpus $N --stepper swe_rk2 --workload tc2_cubed7 --laplacian gmls --steps 3 --warmup 3 > $OUT/bench_swe_cubed7_n$N.json 2> $OUT/bench_swe_cubed7_n$N.err
echo "== swe cubed7 N=$N rc=$?"; cut -c1-500 $OUT/bench_swe_cubed7_n$N.json; tail -3 $OUT/bench_swe_cubed7_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29648 bench.py --gpus $N --stepper swe_rk2 --workload tc2_icos8 --laplacian gmls --steps 1 --warmup 3 > $OUT/bench_swe_icos8_n$N.json 2> $OUT/bench_swe_icos8_n$N.err
echo "== swe icos8 N=$N rc=$?"; cut -c1-500 $OUT/bench_swe_icos8_n$N.json; tail -3 $OUT/bench_swe_icos8_n$N.err
