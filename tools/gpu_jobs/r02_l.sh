#!/bin/bash
# streaming sampler after the round-2 changes: streaming tests on 1 GPU, then the 2-rank parity check
# (tight timeouts: a stall must not eat the lease)
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_batched.py tests/test_gpu_chains200.py tests/test_gpu_sink.py tests/test_gpu_peer.py tests/test_gpu_joint.py tests/test_run_example.py -q -x -m gpu > gpurun_out/r02l_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02l_tests.log
( time GI_CHECK_TIMEOUT=90 timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py ) > gpurun_out/r02l_multi2.log 2>&1
echo "rc=$?" >> gpurun_out/r02l_multi2.log
tail -n 3 gpurun_out/r02l_tests.log; grep "multi_gpu_check\|rc=" gpurun_out/r02l_multi2.log
bash tools/gpu_jobs/r02_n.sh
