#!/bin/bash
# round 2, GPU call C (1 GPU): the reworked bench.py -- mid as a smoke check, then c5
mkdir -p gpurun_out
( time timeout 900 python bench.py --workload mid --steps 20 --warmup 3 ) > gpurun_out/r02c_bench_mid.log 2> gpurun_out/r02c_bench_mid.err
echo "rc=$?" >> gpurun_out/r02c_bench_mid.err
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02c_bench_c5.log 2> gpurun_out/r02c_bench_c5.err
echo "rc=$?" >> gpurun_out/r02c_bench_c5.err
tail -3 gpurun_out/r02c_bench_mid.err gpurun_out/r02c_bench_c5.err
