#!/bin/bash
# GPU check of the head kernels (one GPU): parity tests that touch the head, the head's per-step time and GB/s with
# head_impl = 2 (default) and 1 on the same box, and one ncu --set full capture of head_row2_kernel.
tag=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -x -q -m gpu -k "head or edge_cases or fused_sampling or log_prob_vs" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
CMD="python bench.py --steps 2 --warmup 1 --samples 37888 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/${tag}_impl2.log 2>&1; echo "bench impl2 rc=$?"
$CMD --opt head_impl=1 > gpurun_out/${tag}_impl1.log 2>&1; echo "bench impl1 rc=$?"
python - <<'P'
import json,sys
for t in ("impl2","impl1"):
    import glob
    f=glob.glob("gpurun_out/*_%s.log"%t)
    f=sorted(f)[-1]
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(t, "value", round(d["value"]), "head ms", r["per_class_ms"]["head"], json.dumps(r.get("hbm_kernels"))[:600])
P
ncu --set full --clock-control none --import-source on -k regex:"head_row2" -s 4 -c 2 -f -o gpurun_out/${tag}_head $CMD > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_*
