#!/bin/bash
# 2 GPUs, c5_quarter (2048 rows per GPU = the c5 shard at 8 GPUs): device step vs end-to-end step
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
( time GI_BENCH_WATCHDOG=120 timeout 150 $TR bench.py --gpus 2 --workload c5_quarter --steps 20 --warmup 5 --no-c1 --no-selfcheck ) > gpurun_out/r02n_q2.log 2> gpurun_out/r02n_q2.err
echo "rc=$?" >> gpurun_out/r02n_q2.err
tail -n 2 gpurun_out/r02n_q2.err
