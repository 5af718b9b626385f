#!/bin/sh
# Round 2, 1-GPU call: L2 prefetch A/B of the radix-16 rows, sanitizer test, GPU suite
set -x
OUT=gpurun_out
mkdir -p $OUT
for pf in 0 148 296 592; do
  echo "GFMD_B200_ROWS_PREFETCH=$pf" >> $OUT/r2_c16_prefetch_ab.txt
  GFMD_B200_ROWS_PREFETCH=$pf timeout 200 python tools/stage_times.py 4096 4096 2>&1 | tail -n 1 >> $OUT/r2_c16_prefetch_ab.txt
  GFMD_B200_ROWS_PREFETCH=$pf timeout 200 python tools/stage_times.py 2048 16384 2>&1 | tail -n 1 >> $OUT/r2_c16_prefetch_ab.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2_c16_gpu_suite.txt 2>&1
du -sm $OUT
