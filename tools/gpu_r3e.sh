#!/bin/bash
# r3e (2 GPUs): the pytest multi-GPU tests on the round-end defaults (peer == NCCL exchange bitwise, sharded steppers against
# the oracle with sharded host I/O), log kept under profiles/
TAG=${1:-r3e}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export LPMX_PEER_TIMEOUT_S=20
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -6 | tee $OUT/pytest_multi.log
cp gpurun_out/multi_gpu_check_n*.log gpurun_out/peer_exchange_check_n2.log $OUT/ 2>/dev/null; ls $OUT
