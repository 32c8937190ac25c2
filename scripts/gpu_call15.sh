#!/bin/bash
# round 2, GPU call 15 (1 GPU): RNN sequence kernels (stores after the barrier arrive, prefetched saved tensors): the
# whole tier, then benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c15_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4_flownet_lstm_t50_b16 > gpurun_out/c15_bench_flownet.json 2> gpurun_out/c15_bench_flownet.err
echo done
