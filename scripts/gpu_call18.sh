#!/bin/bash
# round 2, GPU call 18 (2 GPUs): the NCCL world-2 tests with the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c18_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c18_multi.log
echo done
