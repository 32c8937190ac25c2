#!/bin/bash
# round 2, GPU call 19 (1 GPU): vectorised weight packing, folded first layer at W stride 1 (ResNet): whole tier, benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
for wl in cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/c19_bench_$wl.json 2> gpurun_out/c19_bench_$wl.err
done
echo done
