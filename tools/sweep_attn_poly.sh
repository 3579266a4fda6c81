#!/bin/bash
# Sweep of the item-attention exponential split (attn_poly = k of 16 pairs on the FMA pipes, 106 = k 6 with the degree-2
# polynomial) for the lean (v5) and the two-pass (v4) softmax.
# usage (GPU box): bash tools/sweep_attn_poly.sh <tag> "<poly values>" "<lean values>"   -> gpurun_out/<tag>_lean<l>_poly<v>.log
tag=$1
vals=${2:-"0 3 4 5 6 7 8"}
leans=${3:-"1"}
mkdir -p gpurun_out
for l in $leans; do for v in $vals; do
  python bench.py --steps 1 --warmup 1 --samples 37888 --no-cpu-baseline --attn-poly $v --opt attn_lean=$l > gpurun_out/${tag}_lean${l}_poly$v.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_lean${l}_poly$v.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("lean=$l attn_poly=$v", round(d["value"]), "samples/s", "attn_test", round(r["per_class_tflops"]["attn_test"], 1), "TF/s", r["per_class_ms"])
except Exception as e:
    print("lean=$l attn_poly=$v failed", e)
PY
done; done
