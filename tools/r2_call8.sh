#!/bin/sh
# Round 2, N-GPU call: overlapped step without transposes, SM split and chunk count
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
GRID=16384x16384 AB=${AB:-pushes,direct_seq,direct,direct_x8,direct_x12,direct_x24,direct_c4} timeout 600 python -m torch.distributed.run --nnodes=1 \
  --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 tools/stage_times_multi_gpu.py > $OUT/r2_c8_stage_times_16384_${N}gpu.txt 2>&1
du -sm $OUT
