#!/bin/bash
# round 2, GPU call A: the full GPU test suite (new 200-sample chains, near field, fused stress) and
# compute-sanitizer over one tiny invocation of every kernel family
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a_host.log 2>&1
nproc >> gpurun_out/r02a_host.log; free -g >> gpurun_out/r02a_host.log
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r02a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02a_tests.log
python tools/sanitize_workload.py > gpurun_out/r02a_workload.log 2>&1
echo "rc=$?" >> gpurun_out/r02a_workload.log
for tool in memcheck synccheck racecheck; do
  ( time timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_workload.py ) > gpurun_out/r02a_san_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r02a_san_$tool.log
done
tail -5 gpurun_out/r02a_tests.log
