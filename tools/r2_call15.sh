#!/bin/sh
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
GRID=16384x16384 AB=${AB:-direct_40_16_tl,direct_48_24,direct_40_32,direct_56_16,direct_64_16,direct_24_40} timeout 600 python -m torch.distributed.run --nnodes=1 \
  --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29762 tools/stage_times_multi_gpu.py > $OUT/r2_c15_stage_times_16384_${N}gpu.txt 2>&1
du -sm $OUT
