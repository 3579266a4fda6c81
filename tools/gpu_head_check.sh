#!/bin/bash
# GPU check of the head kernels (one GPU): parity tests that touch the head, then the head's per-step time and GB/s for
# each "<engine option>=<value>" given after the tag (e.g. head_impl=1 dec_rows=8192), and with NCU=1 one ncu --set full
# capture of head_row2_kernel.     bash tools/gpu_head_check.sh <tag> [key=value ...]
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -x -q -m gpu -k "head or edge_cases or fused_sampling or log_prob_vs" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
CMD="python bench.py --steps 2 --warmup 1 --samples 37888 --no-cpu-baseline --no-configs"
for opt in default "$@"; do
    o=""; [ "$opt" != default ] && o="--opt $opt"
    $CMD $o > gpurun_out/${tag}_${opt}.log 2>&1; echo "bench $opt rc=$?"
    python - gpurun_out/${tag}_${opt}.log $opt <<'P'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]; h = r["hbm_kernels"]["head"]
        print(sys.argv[2], "value", round(d["value"]), "head ms", r["per_class_ms"]["head"], "GB/s", round(h["achieved"]), "frac", round(h["frac"], 3),
              "gemm ms", r["per_class_ms"]["gemm"], "ms/step", round(d["ms_per_step"], 1))
P
done
if [ -n "$NCU" ]; then
    ncu --set full --clock-control none --import-source on -k regex:"head_row2" -s 4 -c 2 -f -o gpurun_out/${tag}_head $CMD > gpurun_out/${tag}_ncu.log 2>&1
fi
ls -la gpurun_out/${tag}_*
