#!/bin/bash
# round 2, GPU call E (1 GPU): full GPU suite, near-field numbers, wavelet path measurement, and the ncu
# evidence of the bench command (launch list + --set full of the two contractions) and of the wavelet
# kernels.  Every command runs under `timeout`.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 -x ) > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log
timeout 300 python -m pytest tests/test_gpu_nearfield.py -q -s > gpurun_out/r02e_nearfield.log 2>&1
timeout 300 python tools/wavelet_bench.py --workload mid,c2 > gpurun_out/r02e_wavelet.jsonl 2> gpurun_out/r02e_wavelet.err
echo "rc=$?" >> gpurun_out/r02e_wavelet.err
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-selfcheck --no-c1 --min-seconds 0"
$B > gpurun_out/r02e_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv $B > gpurun_out/r02e_ncu_launch.log 2>&1
$B > gpurun_out/r02e_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 8 -c 4 -o gpurun_out/r02e_gemm_c5 $B > gpurun_out/r02e_ncu_full.log 2>&1
W="timeout 300 python tools/wavelet_bench.py --workload mid --reps 3"
$W > gpurun_out/r02e_wplain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwt|spmv" -s 4 -c 6 -o gpurun_out/r02e_wavelet $W > gpurun_out/r02e_ncu_wavelet.log 2>&1
tail -n 4 gpurun_out/r02e_tests.log; tail -n 2 gpurun_out/r02e_ncu_full.log gpurun_out/r02e_ncu_wavelet.log
