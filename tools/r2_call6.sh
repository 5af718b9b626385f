#!/bin/sh
# Round 2, N-GPU call: A/B of the exchange settings on the strong-scaling surface
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
GRID=16384x16384 AB=${AB:-default,chunks1,chunks8,peer_store,direct,sync_nccl,nccl} timeout 600 python -m torch.distributed.run --nnodes=1 \
  --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 tools/stage_times_multi_gpu.py > $OUT/r2_c6_stage_times_16384_${N}gpu.txt 2>&1
du -sm $OUT
