#!/bin/bash
# round 2, GPU call 5: quick regression of the touched tests, bench, then the CTA-pair conv kernel under its own timeout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scan.py -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
( time timeout 180 python -m pytest tests/test_gpu_conv_pair.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/c5_pair.log 2>&1
echo "pair rc=$?" >> gpurun_out/c5_pair.log
nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv >> gpurun_out/c5_pair.log 2>&1
DLIO_CONV_CG2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_cg2.json 2> gpurun_out/c5_bench_cg2.err
echo done
