#!/bin/bash
# r2x (1 GPU): twelve banks in rotation, 2-warp CTAs for small target sets, merged lists: parity, then what a rank of a
# multi-GPU run would see per evaluation at N = 1 .. 8 (tools/rank_size_sweep.py), then the default bench.
TAG=${1:-r2x}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_const_stream.py tests/test_gpu_parity_bve.py -q -m gpu -k 'const or cubed7 or split' 2>&1 | tail -4 | tee $OUT/pytest_const.log
echo "== bitwise"; timeout 300 python tools/pdl_check.py 2>&1 | tail -2 | tee $OUT/pdl_check.txt
echo "== rank sizes"; timeout 900 python tools/rank_size_sweep.py > $OUT/rank_size_sweep.jsonl 2> $OUT/rank_size_sweep.err; tail -3 $OUT/rank_size_sweep.err
python - $OUT/rank_size_sweep.jsonl <<'PY'
import json, sys, collections
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
by = collections.OrderedDict()
for r in rows: by.setdefault(r["n_tgt"], []).append(r)
for n, rs in by.items():
    ring = [r for r in rs if r["config"] == "ring"][0]["ms"]
    print("n_tgt %7d  ring %.3f ms |" % (n, ring), "  ".join("%s %.3f (%.2f)" % (r["config"].replace("banks shape ", ""), r["ms"], ring / r["ms"]) for r in rs if r["config"] != "ring"), "| max diff %.1e" % max(r["rel_diff_vs_ring"] for r in rs))
PY
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-extras > $OUT/bench.json 2> $OUT/bench.err; cut -c1-200 $OUT/bench.json
