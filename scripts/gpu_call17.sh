#!/bin/bash
# round 2, GPU call 17 (1 GPU): smoke() with its second (default-path) step, then the sanitizer target
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c17_smoke.log 2>&1
SANITIZE_TIMEOUT=650 bash scripts/sanitize.sh > gpurun_out/c17_sanitize.log 2>&1
echo "sanitize rc=$?" >> gpurun_out/c17_sanitize.log
echo done
