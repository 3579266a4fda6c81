#!/bin/bash
# timing sweep of attn_tc6 variants (GPU box):  bash tools/sweep_v6.sh <tag> "<lib suffixes>" "<poly values>" "<extra --opt k=v ...>"
tag=$1; libs=${2:-"_"}; polys=${3:-"5"}; extra=${4:-""}
for l in $libs; do for k in $polys; do
  lib=tools/_bin/lib6${l#_}.so
  NPE_PFN_B200_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 1 --samples 37888 --no-configs --no-cpu-baseline --opt attn_impl=2 --attn-poly $k $extra > gpurun_out/${tag}_${l}_$k.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${l}_$k.log").read().strip().splitlines()[-1]); r = d["roofline"]
    print("lib6${l#_} poly $k $extra:", round(d["value"]), "samples/s  attn_test", round(r["per_class_tflops"]["attn_test"], 1), "TF/s", r["per_class_ms"]["attn_test"], "ms  clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("lib6${l#_} poly $k failed", e)
PY
done; done
