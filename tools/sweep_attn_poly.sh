#!/bin/bash
# Sweep of the item-attention exponential split (attn_poly = k of 16 pairs on the FMA pipes, +100 = degree-2 polynomial).
# usage (GPU box): bash tools/sweep_attn_poly.sh <tag> "<values>"   -> gpurun_out/<tag>_poly<v>.log
tag=$1; shift
vals=${1:-"0 4 5 6 7 8 10 106 107 108"}
mkdir -p gpurun_out
for v in $vals; do
  python bench.py --steps 1 --warmup 1 --samples 37888 --no-cpu-baseline --attn-poly $v > gpurun_out/${tag}_poly$v.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_poly$v.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("attn_poly=$v", round(d["value"]), "samples/s", "attn_test", round(r["per_class_tflops"]["attn_test"], 1), "TF/s", r["per_class_ms"])
except Exception as e:
    print("attn_poly=$v failed", e)
PY
done
