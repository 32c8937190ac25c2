#!/bin/bash
# round 2, GPU call 31 (1 GPU): the added glue tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c31_glue.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c31_glue.log
echo done
