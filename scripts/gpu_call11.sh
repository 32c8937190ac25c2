#!/bin/bash
# round 2, GPU call 11 (1 GPU): quick regression of the touched tests, PointSeg / headline bench, then the ncu evidence
# of the final kernels: launch list of one eager step (time, DRAM bytes, issue activity) and two --set full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2_pointseg_lstm_b32 > gpurun_out/c11_bench_pointseg.json 2> gpurun_out/c11_bench_pointseg.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 800 --csv \
    --log-file gpurun_out/c11_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c11_ncu_bench.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2_kernel|conv_tc_kernel|bn_pool3_fwd_tma' -c 14 -o gpurun_out/c11_full_fwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c11_ncu_full_fwd.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc|bwd_apply_tma|pool_bwd_sums' -c 16 -o gpurun_out/c11_full_bwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c11_ncu_full_bwd.log 2>&1
ls -la gpurun_out/c11_full_* >> gpurun_out/c11_ncu_full_bwd.log
echo done
