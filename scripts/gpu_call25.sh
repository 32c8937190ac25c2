#!/bin/bash
# round 2, GPU call 25 (1 GPU): bias gradients from the weight-gradient launch: whole tier, smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c25_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c25_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c25_smoke.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg4_flownet_lstm_t50_b16 > gpurun_out/c25_bench_flownet.json 2> gpurun_out/c25_bench_flownet.err
echo done
