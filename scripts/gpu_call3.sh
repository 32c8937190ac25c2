#!/bin/bash
# round 2, GPU call 3: test tier with the bulk-copy pooling kernels, bench A/B (pool_tma on/off), ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_tma.json 2> gpurun_out/c3_bench_tma.err
DLIO_POOL_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_notma.json 2> gpurun_out/c3_bench_notma.err
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/c3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c3_ncu_bench.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -c 12 -o gpurun_out/c3_pool_tma \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c3_ncu_full.log 2>&1
echo done
