#!/bin/bash
# round 2, GPU call 16 (8 GPUs): where the N = 8 step time goes: no exchange at all, one all-reduce after a single
# graph, the split backward (default), and the default with NCCL's CTA count capped
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
port=29700
run() {  # name, env...
  name=$1; shift
  port=$((port + 1))
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c16_n8_$name.json 2> gpurun_out/c16_n8_$name.err
}
run noexchange DLIO_NO_EXCHANGE=1
run single_graph DLIO_SPLIT_BWD=0
run split DLIO_SPLIT_BWD=1
run split_maxctas8 DLIO_SPLIT_BWD=1 NCCL_MAX_CTAS=8
run split_maxctas32 DLIO_SPLIT_BWD=1 NCCL_MAX_CTAS=32 NCCL_MIN_CTAS=32
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29790 \
    bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | grep -E "NVLS|Channel|nChannels|algo|Connected|threadThresholds" | head -40 > gpurun_out/c16_nccl_info.txt
echo done
