#!/bin/bash
# round 2, final 1-GPU call: the whole GPU suite, smoke(), then the c5 bench for the record
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/r02o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02o_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r02o_smoke.log
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02o_bench_c5.log 2> gpurun_out/r02o_bench_c5.err
echo "rc=$?" >> gpurun_out/r02o_bench_c5.err
tail -n 3 gpurun_out/r02o_tests.log gpurun_out/r02o_smoke.log gpurun_out/r02o_bench_c5.err
