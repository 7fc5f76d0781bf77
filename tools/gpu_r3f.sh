#!/bin/bash
# r3f (1 GPU): BASELINE configs[4]'s upper sizes on the pipelined bank path: synthetic N = 3e6 (one RK4 step) and 1e7 (one
# velocity evaluation = 1e14 interactions)
TAG=${1:-r3f}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 280 python tools/synthetic_sweep.py --sizes 1e7 --n-check 8 > $OUT/synthetic_sweep_n1.jsonl 2> $OUT/synthetic.err; cut -c1-500 $OUT/synthetic_sweep_n1.jsonl; tail -2 $OUT/synthetic.err
