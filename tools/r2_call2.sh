#!/bin/sh
# Round 2, multi-GPU call:  gpurun --gpus N --timeout 900 -- 'sh tools/r2_call2.sh'
set -x
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
nvidia-smi topo -m > $OUT/r2_c2_topo_$N.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $OUT/r2_c2_multi_gpu_tests_$N.txt 2>&1
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
AB=${AB_WEAK:-default,sync_nccl,peer_store,direct,chunks8,chunks2,chunks1,rows_r8} timeout 600 $RUN --master-port 29711 \
  tools/stage_times_multi_gpu.py > $OUT/r2_c2_stage_times_weak_${N}gpu.txt 2>&1
GRID=16384x16384 AB=${AB_STRONG:-default,sync_nccl,peer_store,direct,chunks8} timeout 600 $RUN --master-port 29712 \
  tools/stage_times_multi_gpu.py > $OUT/r2_c2_stage_times_16384_${N}gpu.txt 2>&1
du -sm $OUT
