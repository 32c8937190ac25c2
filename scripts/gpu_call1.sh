#!/bin/bash
# round 2, GPU call 1: full GPU test tier (all failures, not -x), bench A/Bs, launch list with DRAM bytes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench_base.json 2> gpurun_out/c1_bench_base.err
DLIO_ENC_STREAMS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_streams.json 2> gpurun_out/c1_bench_streams.err
DLIO_ENC_STREAMS=1 DLIO_EW_BLOCK=128 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_streams_b128.json 2> gpurun_out/c1_bench_streams_b128.err
DLIO_SKIP_DZ=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_nodzskip.json 2> gpurun_out/c1_bench_nodzskip.err
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c1_ncu_bench.log 2>&1
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err
echo done
