#!/bin/sh
# Round 2, N-GPU call: the bench as the driver runs it under torchrun
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29751 \
  bench.py --gpus $N --steps 30 --warmup 5 > $OUT/r2_c11_bench_${N}gpu.json 2> $OUT/r2_c11_bench_${N}gpu.err
du -sm $OUT
