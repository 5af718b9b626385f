#!/bin/sh
# Round 2, N-GPU call: slab parity test at N ranks (default exchange) and the bench as the driver runs it
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python -m pytest "tests/test_multi_gpu.py::test_slab_parity[$N]" -m gpu -q -x > $OUT/r2_c5_slab_parity_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
  bench.py --gpus $N --steps 30 --warmup 5 > $OUT/r2_c5_bench_${N}gpu.json 2> $OUT/r2_c5_bench_${N}gpu.err
du -sm $OUT
