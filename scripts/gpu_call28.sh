#!/bin/bash
# round 2, GPU call 28 (1 GPU): --set full capture of the whole-sequence RNN kernels (source-level stall analysis)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DLIO_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rnn_seq' -c 4 -o gpurun_out/c28_rnn \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg4_flownet_lstm_t50_b16 > gpurun_out/c28_ncu.log 2>&1
ls -la gpurun_out/ >> gpurun_out/c28_ncu.log
echo done
