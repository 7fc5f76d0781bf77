#!/bin/bash
# r2m (8 GPUs): contract bench with both exchanges, peer-exchange bit parity at N = 8, icos-8 at N = 8, synthetic sweep up to 1e7
TAG=${1:-r2m}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/gpu_r2l.sh $TAG $N
export LPMX_PEER_TIMEOUT_S=10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/peer_exchange_check.py --time cubed-7 --steps 8 --skip-oracle > $OUT/peer_check_n$N.txt 2> $OUT/peer_check_n$N.err
echo "peer_exchange_check exit $?"; grep -c "bitwise: True" $OUT/peer_check_n$N.txt; tail -3 $OUT/peer_check_n$N.txt
LPMX_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 \
  bench.py --gpus $N --workload gauss_icos8 --steps 3 --warmup 3 --no-extras > $OUT/bench_icos8_n$N.json 2> $OUT/bench_icos8_n$N.err
echo "== icos-8 N=$N: $(python -c "import json; d=json.loads(open('$OUT/bench_icos8_n$N.json').read().strip().splitlines()[-1]); print('%.4e inter/s %.1f ms e2e %.1f ms parity %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'].get('max_rel_err')))" 2>&1)"
LPMX_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29630 \
  tools/synthetic_sweep.py --sizes 1e6,3e6,1e7 --steps 1 > $OUT/synthetic_sweep_n$N.jsonl 2> $OUT/synthetic_sweep_n$N.err
echo "== synthetic sweep"; cat $OUT/synthetic_sweep_n$N.jsonl | cut -c1-330
