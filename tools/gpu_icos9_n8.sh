#!/bin/bash
# 8-GPU visit: BASELINE configs[2] -- sphere_gaussian_vortex on icosTriSphereSeed depth 9 (9.6 M targets x 5.24 M leaf sources),
# BVERK4 step, target-sharded over 8 B200 (gpurun --gpus 8).  One timed step after three warm-up steps: a step is ~15 s.
TAG=${1:-r1x}; N=8
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29628 bench.py --gpus $N --workload gauss_icos9 --steps 1 --warmup 3 > $OUT/bench_icos9_n$N.json 2> $OUT/bench_icos9_n$N.err
echo "== bench icos9 N=$N rc=$?"; cut -c1-600 $OUT/bench_icos9_n$N.json; tail -3 $OUT/bench_icos9_n$N.err
