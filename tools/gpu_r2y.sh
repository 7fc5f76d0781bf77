#!/bin/bash
# r2y (1 GPU): what a rank of a multi-GPU run would see per evaluation at N = 1 .. 8 (tools/rank_size_sweep.py)
TAG=${1:-r2y}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== rank sizes"; timeout 900 python tools/rank_size_sweep.py > $OUT/rank_size_sweep.jsonl 2> $OUT/rank_size_sweep.err; tail -3 $OUT/rank_size_sweep.err
python - $OUT/rank_size_sweep.jsonl <<'PY'
import json, sys, collections
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
by = collections.OrderedDict()
for r in rows: by.setdefault(r["n_tgt"], []).append(r)
for n, rs in by.items():
    ring = [r for r in rs if r["config"] == "ring"][0]["ms"]
    print("n_tgt %7d  ring %.3f ms |" % (n, ring), "  ".join("%s %.3f (%.2f)" % (r["config"].replace("batch ", "b"), r["ms"], ring / r["ms"]) for r in rs if r["config"] != "ring"), "| max diff %.1e" % max(r["rel_diff_vs_ring"] for r in rs))
PY
