#!/bin/bash
# round 2, GPU call 33 (1 GPU): last check of the final tree -- whole tier, smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider ) > gpurun_out/c33_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c33_pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c33_smoke.log 2>&1
echo done
