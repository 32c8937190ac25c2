#!/bin/bash
# round 2, GPU call 8 (1 GPU): single-pass backward A/B (gradient error + step time), the secondary workloads at
# their per-GPU batch with the per-class breakdown, ncu launch list of one PointSeg step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/ab_single_pass.py > gpurun_out/c8_ab_single_pass.json 2> gpurun_out/c8_ab_single_pass.err
DLIO_BWD_SINGLE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_single.json 2> gpurun_out/c8_bench_single.err
for wl in cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/c8_bench_$wl.json 2> gpurun_out/c8_bench_$wl.err
done
DLIO_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 3000 --csv \
    --log-file gpurun_out/c8_launches_pointseg.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c8_ncu_pointseg.log 2>&1
echo done
