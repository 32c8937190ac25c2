#!/bin/bash
# round 2, GPU call 14 (8 GPUs): the headline workload and the three BASELINE.json 8-GPU configurations at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c14_gpus.txt 2>&1
port=29600
for wl in "" cfg2_pointseg_lstm_b32 cfg3_resnet_gru_b64 cfg4_flownet_lstm_t50_b16; do
  port=$((port + 1))
  name=${wl:-headline}
  extra=""
  [ -n "$wl" ] && extra="--workload $wl"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --steps 10 --warmup 3 $extra > gpurun_out/c14_bench_n8_$name.json 2> gpurun_out/c14_bench_n8_$name.err
done
port=$((port + 1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/c14_bench_n4_headline.json 2> gpurun_out/c14_bench_n4_headline.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_n1_headline.json 2> gpurun_out/c14_bench_n1_headline.err
echo done
