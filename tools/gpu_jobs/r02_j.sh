#!/bin/bash
# round 2, GPU call J (N GPUs): bench.py c5 at N = $1 with the peer-memory exchange
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
( time GI_BENCH_WATCHDOG=500 timeout 700 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r02j_bench_n$N.log 2> gpurun_out/r02j_bench_n$N.err
echo "rc=$?" >> gpurun_out/r02j_bench_n$N.err
tail -n 3 gpurun_out/r02j_bench_n$N.err
