#!/bin/bash
# round 2, GPU call D (8 GPUs): 4-rank parity check of the exchanges, then bench.py c5 at N = 8 with the
# peer-memory exchange (default) and with the NCCL all-reduce hooks for comparison
mkdir -p gpurun_out
( time GI_CHECK_TIMEOUT=300 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py ) > gpurun_out/r02d_multi4.log 2>&1
echo "rc=$?" >> gpurun_out/r02d_multi4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
( time GI_BENCH_WATCHDOG=600 timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02d_bench_n8_peer.log 2> gpurun_out/r02d_bench_n8_peer.err
echo "rc=$?" >> gpurun_out/r02d_bench_n8_peer.err
( time GI_SHARD_EXCHANGE=nccl GI_BENCH_WATCHDOG=600 timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-c1 --no-selfcheck ) > gpurun_out/r02d_bench_n8_nccl.log 2> gpurun_out/r02d_bench_n8_nccl.err
echo "rc=$?" >> gpurun_out/r02d_bench_n8_nccl.err
tail -n 4 gpurun_out/r02d_multi4.log; tail -n 3 gpurun_out/r02d_bench_n8_peer.err; tail -n 3 gpurun_out/r02d_bench_n8_nccl.err
