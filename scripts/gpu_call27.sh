#!/bin/bash
# round 2, GPU call 27 (1 GPU): final evidence -- default bench line, launch list of one eager step, --set full captures
# (8 launches each: gpurun brings back at most 64 MiB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err
for wl in cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  timeout 400 python bench.py --no-cpu-baseline --workload $wl > gpurun_out/c27_bench_$wl.json 2> gpurun_out/c27_bench_$wl.err
done
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 800 --csv \
    --log-file gpurun_out/c27_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c27_ncu_bench.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2_kernel|conv_tc_kernel|bn_pool3_fwd_tma|bn_apply_rows' -c 8 -o gpurun_out/c27_full_fwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c27_ncu_full_fwd.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc|bwd_apply_tma|pool_bwd_sums|bn_bwd_reduce_flat' -c 8 -o gpurun_out/c27_full_bwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c27_ncu_full_bwd.log 2>&1
echo done
