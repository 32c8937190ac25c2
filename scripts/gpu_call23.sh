#!/bin/bash
# round 2, GPU call 23 (1 GPU): 64-channel output tiles on the CTA-pair kernel: pair tests under their own timeout, A/B
# benches (headline, ResNet), then the model tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_conv_pair.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c23_pair.log 2>&1
echo "pair rc=$?" >> gpurun_out/c23_pair.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
DLIO_CG2_N64=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c23_bench_off.json 2> gpurun_out/c23_bench_off.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg3_resnet_gru_b64 > gpurun_out/c23_bench_resnet.json 2> gpurun_out/c23_bench_resnet.err
DLIO_CG2_N64=0 timeout 400 python bench.py --no-cpu-baseline --workload cfg3_resnet_gru_b64 > gpurun_out/c23_bench_resnet_off.json 2> gpurun_out/c23_bench_resnet_off.err
( time timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/c23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c23_pytest.log
echo done
