#!/bin/bash
# round 2, GPU call 9 (1 GPU): folded-split first layer (fp16) under its own timeout, graph split test, the model /
# full-size parity tests, bench A/B (DLIO_FIRST_F16), PointSeg bench (squeeze without fp32 copy, zero_tail)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 300 python -m pytest "tests/test_gpu_ops.py::test_first_layer_space_to_depth_at_64x2048" tests/test_gpu_graph.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c9_first.log 2>&1
echo "first rc=$?" >> gpurun_out/c9_first.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
DLIO_FIRST_F16=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_tf32first.json 2> gpurun_out/c9_bench_tf32first.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c9_bench_pointseg.json 2> gpurun_out/c9_bench_pointseg.err
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
echo done
