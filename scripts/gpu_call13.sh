#!/bin/bash
# round 2, GPU call 13 (1 GPU): flat BN-backward reduce kernel: op tests, model tests, benches, launch lists of the
# secondary workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/c13_ops.log 2>&1
echo "ops rc=$?" >> gpurun_out/c13_ops.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
for wl in cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/c13_bench_$wl.json 2> gpurun_out/c13_bench_$wl.err
  DLIO_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 3000 --csv \
    --log-file gpurun_out/c13_launches_$wl.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload $wl > gpurun_out/c13_ncu_$wl.log 2>&1
done
echo done
