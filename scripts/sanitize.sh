#!/bin/bash
# compute-sanitizer target (run on a GPU box): memcheck and initcheck over the per-op C-ABI tests and one full train
# step of the smallest golden model.  Writes gpurun_out/sanitize_{memcheck,initcheck}.log; exit code != 0 if either
# tool reports an error.
# usage: gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_conv_pair.py tests/test_gpu_glue.py tests/test_gpu_scan.py tests/test_gpu_ops.py"
LIMIT=${SANITIZE_TIMEOUT:-1200}
rc=0
for tool in memcheck initcheck; do
  timeout $LIMIT $SAN --tool $tool --error-exitcode 9 --launch-timeout 0 \
      python -m pytest $TESTS -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  r=$?
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
  [ $r -ne 0 ] && rc=$r
done
exit $rc
