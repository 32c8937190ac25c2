#!/bin/bash
# round 2, GPU call 2: full GPU test tier after the test / kernel fixes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.log 2>&1
echo done
