#!/bin/bash
# usage: tools/sweep.sh tag variant1 variant2 ...   (run on the GPU box; "main" = libb2ndt.so)
tag=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset B2NDT_LIB; else export B2NDT_LIB=$PWD/lidar_slam_b200/_lib/libb2ndt_$v.so; fi
  timeout 200 python bench.py --steps 3 --warmup 3 --cpu-sample 4 --raw-frames 0 > gpurun_out/sweep_${tag}_$v.json 2> gpurun_out/sweep_${tag}_$v.err
  echo "$v exit=$?"
  grep "ndt batch timing" gpurun_out/sweep_${tag}_$v.err | sed -n '8p'
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${tag}_$v.json"))
    print("  $v value %.0f ms/step %.2f e2e %.0f single %s c5 %.2f ms parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["single_match_ms"], d["config5"]["wall_ms_all"], d["parity"]))
except Exception as e:
    print("  $v failed", e)
PY
done
