#!/bin/bash
# round 2, GPU call 26 (2 GPUs): final code through torchrun at N = 2 (the driver's launch line), NCCL tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c26_bench_n2.json 2> gpurun_out/c26_bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/c26_bench_n2_reference.json 2> gpurun_out/c26_bench_n2_reference.err
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c26_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c26_multi.log
echo done
