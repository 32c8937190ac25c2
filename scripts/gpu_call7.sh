#!/bin/bash
# round 2, GPU call 7 (1 GPU): wgrad pair kernel under its own timeout, the whole -m gpu tier, smoke(), bench with the
# wgrad pair kernel on / off, the ncu launch list of one eager step (times + DRAM bytes per launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 240 python -m pytest tests/test_gpu_conv_pair.py -m gpu -q -p no:cacheprovider ) > gpurun_out/c7_pair.log 2>&1
echo "pair rc=$?" >> gpurun_out/c7_pair.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
DLIO_WGRAD_CG2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench_nowg2.json 2> gpurun_out/c7_bench_nowg2.err
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c7_smoke.log 2>&1
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 800 --csv \
    --log-file gpurun_out/c7_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c7_ncu_bench.log 2>&1
echo done
