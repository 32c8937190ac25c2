#!/bin/bash
# round 2, GPU call 4: op / scan / protocol tests, bench, ncu of the leaner pooling kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scan.py tests/test_gpu_train_protocol.py tests/test_gpu_glue.py tests/test_gpu_model.py -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 800 --csv \
    --log-file gpurun_out/c4_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c4_ncu_bench.log 2>&1
echo done
