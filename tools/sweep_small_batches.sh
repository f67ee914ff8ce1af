#!/bin/bash
# usage (GPU box): tools/sweep_small_batches.sh "500 250 128" "0 1 2 4 8"   -- bench.py ms/step per batch size and cluster width
for f in ${1:-500 250 128}; do for c in ${2:-0 1 2 4 8}; do
  timeout 160 python bench.py --frames $f --batch-cluster $c --steps 5 --warmup 3 --cpu-sample 2 --raw-frames 0 --hypotheses 128 2>/dev/null > gpurun_out/ss_${f}_${c}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ss_${f}_${c}.json")); print("frames $f cluster $c: ms/step %.3f  matches/s %.0f  c5(128) %.3f ms" % (d["ms_per_step"], d["value"], d["config5"]["wall_ms_all"]))
except Exception as e: print("frames $f cluster $c failed", e)
PY
done; done
