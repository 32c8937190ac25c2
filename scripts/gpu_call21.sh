#!/bin/bash
# round 2, GPU call 21 (1 GPU): final state -- the whole tier, smoke(), the default bench line (as the driver runs it), the
# reference arm, the secondary workloads, the ncu launch list of one eager step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_diag.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider ) > gpurun_out/c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c21_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c21_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c21_bench_20.json 2> gpurun_out/c21_bench_20.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c21_bench_reference.json 2> gpurun_out/c21_bench_reference.err
for wl in cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/c21_bench_$wl.json 2> gpurun_out/c21_bench_$wl.err
done
DLIO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 800 --csv \
    --log-file gpurun_out/c21_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c21_ncu_bench.log 2>&1
echo done
