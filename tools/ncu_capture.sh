#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box, one GPU):  bash tools/ncu_capture.sh <tag>
#   1. plain run (must exit 0 before anything is profiled)
#   2. launch list: per-launch duration + DRAM bytes of 1400 launches from the middle of the step
#   3. one --set full capture of the item-attention kernel, of five consecutive projection / fused-MLP launches, of the
#      encoder / K/V cache writer / compaction kernels and of two head launches
tag=$1
CMD="python bench.py --steps 1 --warmup 1 --samples 16384 --no-cpu-baseline --no-configs"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 1400 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 140 -c 1 -f -o gpurun_out/${tag}_attn_tc $CMD > gpurun_out/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|mlp_tc" -s 500 -c 5 -f -o gpurun_out/${tag}_gemm_tc $CMD > gpurun_out/${tag}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_kernel|kv_cache_kernel|compact_append" -s 20 -c 4 -f -o gpurun_out/${tag}_hbm $CMD > gpurun_out/${tag}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"head_row" -s 2 -c 2 -f -o gpurun_out/${tag}_head $CMD > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out/${tag}_*
