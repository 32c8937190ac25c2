#!/bin/bash
# round 2, GPU call 32 (2 GPUs): the odometry-net all-reduce captured into graph A: NCCL tests under a timeout, then
# N = 2 bench with DLIO_ODOM_AR_IN_GRAPH on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c32_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c32_multi.log
DLIO_ODOM_AR_IN_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c32_bench_n2_odom.json 2> gpurun_out/c32_bench_n2_odom.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29822 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c32_bench_n2_split.json 2> gpurun_out/c32_bench_n2_split.err
DLIO_NO_EXCHANGE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29823 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c32_bench_n2_noex.json 2> gpurun_out/c32_bench_n2_noex.err
echo done
