#!/bin/bash
# round 2, GPU call 20 (1 GPU): backward of narrow layers (Fire squeezes) on the tensor cores: whole tier, PointSeg bench
# with the switch on / off, headline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c20_bench_pointseg.json 2> gpurun_out/c20_bench_pointseg.err
DLIO_NARROW_TC_BWD=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c20_bench_pointseg_simt.json 2> gpurun_out/c20_bench_pointseg_simt.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
echo done
