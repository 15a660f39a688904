#!/bin/bash
# round 2, GPU call B (2 GPUs): the peer-memory exchange -- single-rank tests, then the row-sharded
# parity check over 2 ranks (peer memory, NCCL hooks and host-driven exchanges against the oracle)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_batched.py tests/test_gpu_sink.py -x -q ) > gpurun_out/r02b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_tests.log
( time GI_CHECK_TIMEOUT=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py ) > gpurun_out/r02b_multi2.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_multi2.log
tail -5 gpurun_out/r02b_tests.log; tail -15 gpurun_out/r02b_multi2.log
