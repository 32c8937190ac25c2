#!/bin/bash
# round 2, GPU call 6 (2 GPUs): wgrad pair kernel tests, NCCL world-2 tests, bench at N = 1 (wgrad pairs on / off) and
# N = 2 (split backward on / off) on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_conv_pair.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c6_pair.log 2>&1
echo "pair rc=$?" >> gpurun_out/c6_pair.log
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err
DLIO_WGRAD_CG2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_n1_nowg2.json 2> gpurun_out/c6_bench_n1_nowg2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
DLIO_SPLIT_BWD=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c6_bench_n2_nosplit.json 2> gpurun_out/c6_bench_n2_nosplit.err
echo done
