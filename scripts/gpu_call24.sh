#!/bin/bash
# round 2, GPU call 24 (1 GPU): fewer blocks in the reduction passes that end in fp64 atomics; pair test fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_conv_pair.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/c24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c24_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c24_bench_pointseg.json 2> gpurun_out/c24_bench_pointseg.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg3_resnet_gru_b64 > gpurun_out/c24_bench_resnet.json 2> gpurun_out/c24_bench_resnet.err
echo done
