#!/bin/bash
# round 2, GPU call I (1 GPU): the c5 bench for the record, then ncu --set full of gemm_adj alone at c5
# (a second profiled kernel in one process returned nan counters), the v3 single-pass kernel, the CSR matvec
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02i_bench_c5.log 2> gpurun_out/r02i_bench_c5.err
echo "rc=$?" >> gpurun_out/r02i_bench_c5.err
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-selfcheck --no-c1 --min-seconds 0"
$B > gpurun_out/r02i_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_adj -s 4 -c 1 -o gpurun_out/r02i_gemm_adj_c5 $B > gpurun_out/r02i_ncu_adj.log 2>&1
F="timeout 300 python tools/fused_bench.py --rows 4096 --reps 2"
$F > gpurun_out/r02i_fplain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_pass -s 2 -c 1 -o gpurun_out/r02i_fused_v3 $F > gpurun_out/r02i_ncu_fused.log 2>&1
W="timeout 300 python tools/wavelet_bench.py --workload mid --reps 3"
$W > gpurun_out/r02i_wplain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv -s 2 -c 2 -o gpurun_out/r02i_spmv $W > gpurun_out/r02i_ncu_spmv.log 2>&1
tail -n 3 gpurun_out/r02i_bench_c5.err gpurun_out/r02i_ncu_adj.log gpurun_out/r02i_ncu_fused.log gpurun_out/r02i_ncu_spmv.log
