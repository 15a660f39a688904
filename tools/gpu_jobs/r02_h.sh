#!/bin/bash
# ncu --set full of the fused single-pass kernel (source-level stall reasons)
mkdir -p gpurun_out
F="timeout 300 python tools/fused_bench.py --rows 4096 --reps 2"
$F > gpurun_out/r02h_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_pass -s 2 -c 1 -o gpurun_out/r02h_fused $F > gpurun_out/r02h_ncu.log 2>&1
tail -n 3 gpurun_out/r02h_ncu.log
