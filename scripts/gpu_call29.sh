#!/bin/bash
# round 2, GPU call 29 (1 GPU): RNN forward kernel with the per-thread row count as a compile-time constant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_train_protocol.py -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/c29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c29_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg4_flownet_lstm_t50_b16 > gpurun_out/c29_bench_flownet.json 2> gpurun_out/c29_bench_flownet.err
echo done
