#!/bin/sh
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29761 \
  bench.py --gpus $N --steps 30 --warmup 5 > $OUT/r2_c14_bench_${N}gpu.json 2> $OUT/r2_c14_bench_${N}gpu.err
GRID=16384x16384 AB=${AB:-direct_tl,direct_40_16} timeout 600 python -m torch.distributed.run --nnodes=1 \
  --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29762 tools/stage_times_multi_gpu.py > $OUT/r2_c14_stage_times_16384_${N}gpu.txt 2>&1
du -sm $OUT
