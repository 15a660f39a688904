#!/bin/bash
# fused single-pass kernel v2 (retained row in registers): parity tests + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r02g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02g_tests.log
timeout 300 python tools/fused_bench.py --rows 4096 --reps 10 > gpurun_out/r02g_fused_4096.log 2>&1
timeout 300 python tools/fused_bench.py --rows 8192 --reps 5 > gpurun_out/r02g_fused_16384.log 2>&1
tail -n 3 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_fused_4096.log gpurun_out/r02g_fused_16384.log | tail -n 12
