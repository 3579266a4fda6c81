# Tuning only: throughput of the item-attention variants (attn_lean 1 / 2) when the item Q/K projection weights are scaled up
# (sharper scores, rows change their reference maximum often).  usage (GPU box): bash tools/sweep_item_gain.sh
for g in 1 2.5 6; do for l in 1 2; do
python bench.py --steps 1 --warmup 1 --samples 37888 --no-cpu-baseline --item-gain $g --opt attn_lean=$l --opt attn_debug=1 > gpurun_out/r56_g${g}_l$l.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r56_g${g}_l$l.log").read().strip().splitlines()[-1]); r = d["roofline"]
print("gain=$g lean=$l", round(d["value"]), "samples/s attn_test", round(r["per_class_tflops"]["attn_test"], 1), "TF/s")
PY
done; done
