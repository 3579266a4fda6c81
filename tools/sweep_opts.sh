#!/bin/bash
# one short bench per engine-option set (GPU box):  bash tools/sweep_opts.sh <tag> "k=v k=v" "k=v" ...   ("-" = defaults)
tag=$1; shift
mkdir -p gpurun_out
i=0
for set in "$@"; do
    o=""
    if [ "$set" != "-" ]; then for kv in $set; do o="$o --opt $kv"; done; fi
    python bench.py --steps 2 --warmup 1 --samples 37888 --no-cpu-baseline --no-configs $o > gpurun_out/${tag}_$i.log 2>&1
    python - gpurun_out/${tag}_$i.log "$set" <<'P'
import json, sys
ok = False
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]; ok = True
        print(f"{sys.argv[2]:40s} value {d['value']:8.0f}  attn_test {r['per_class_tflops']['attn_test']:6.1f} TF  attn_ctx {r['per_class_tflops']['attn_ctx']:6.1f} TF  "
              f"ms/step {d['ms_per_step']:7.1f}  gemm {r['per_class_ms']['gemm']:6.1f} ms  mlp {r['per_class_ms']['mlp']:5.1f} ms  sm {d['clocks']['sm_mhz']}")
if not ok:
    print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-800:])
P
    i=$((i+1))
done
