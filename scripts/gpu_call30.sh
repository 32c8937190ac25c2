#!/bin/bash
# round 2, GPU call 30 (1 GPU): the final state once more -- whole tier, smoke(), default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c30_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c30_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c30_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err
echo done
